"""Synthetic corpora shaped like the datasets BASELINE.json names (no network, no real data).

Generator: low-intrinsic-dimension latent model  x = z·A + noise·eps,
z ~ N(0, I_rank), A ~ N(0,1)^{rank x dim}; base rows and queries come from the same
stream.  Isotropic Gaussians give an unrealistically hard graph (recall@10 ~0.58 at
ef=100 on 1M x 128); the latent model behaves like SIFT/DEEP descriptors.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (dim, metric, M)      metric 0 = L2, 1 = inner product
    "sift": (128, 0, 16),
    "gist": (960, 0, 32),
    "cohere": (768, 1, 32),
    "deep": (96, 0, 16),
    "msturing": (96, 0, 16),
}


def latent_gaussian(n: int, dim: int, *, rank: int = 16, noise: float = 0.1, seed: int = 1,
                    normalize: bool = False, chunk: int = 1 << 16, stream: int = 0) -> np.ndarray:
    """n x dim float32 rows of the latent model; deterministic in (row, dim, rank, noise, seed, stream)."""
    A = np.random.default_rng(seed).standard_normal((rank, dim)).astype(np.float32)
    rng = np.random.default_rng([seed, stream, 0x5eed])
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, chunk):          # always draw whole chunks: row i does not depend on n
        e = min(n, s + chunk)
        z = rng.standard_normal((chunk, rank), dtype=np.float32)
        eps = rng.standard_normal((chunk, dim), dtype=np.float32)
        out[s:e] = (z[: e - s] @ A) + noise * eps[: e - s]
    if normalize:
        out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def latent_gaussian_rows(lo: int, hi: int, dim: int, *, rank: int = 16, noise: float = 0.1, seed: int = 1,
                         normalize: bool = False, stream: int = 0, chunk: int = 1 << 16) -> np.ndarray:
    """Rows [lo, hi) of a latent-model corpus whose rows are addressable: chunk c (rows c*chunk ..) has its
    own generator, so a rank of a sharded run draws only its shard (the 100M-shaped corpora never exist in
    one piece).  Same model as latent_gaussian, different stream of random numbers."""
    A = np.random.default_rng(seed).standard_normal((rank, dim)).astype(np.float32)
    out = np.empty((hi - lo, dim), dtype=np.float32)
    for c in range(lo // chunk, (hi + chunk - 1) // chunk):
        rng = np.random.default_rng([seed, stream, 0xC0DE, c])
        z = rng.standard_normal((chunk, rank), dtype=np.float32)
        eps = rng.standard_normal((chunk, dim), dtype=np.float32)
        rows = z @ A
        rows += noise * eps
        a, b = max(lo, c * chunk), min(hi, (c + 1) * chunk)
        out[a - lo: b - lo] = rows[a - c * chunk: b - c * chunk]
    if normalize:
        out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def make_dataset(n: int, nq: int, dim: int, *, metric: int = 0, rank: int = 16, noise: float = 0.1,
                 seed: int = 1):
    """(base[n,dim], queries[nq,dim]) from the same distribution (same mixing matrix A, disjoint
    streams); rows L2-normalised for the IP metric.  base[i] depends only on (i, dim, rank, noise, seed)."""
    base = latent_gaussian(n, dim, rank=rank, noise=noise, seed=seed, normalize=(metric == 1))
    qry = latent_gaussian(nq, dim, rank=rank, noise=noise, seed=seed, normalize=(metric == 1), stream=1)
    return base, qry
