"""Golden fixtures for the two "next" strategies, generated from the REFERENCE ITSELF
(oracle/_ref/libhsref_hnsw_*.so = the unmodified hnswalg.h / hnswalg_slimzero.h):

    python tests/golden/make_golden_hnsw.py

  hnsw_l2_1k.graph       HierarchicalNSW::saveIndex for 1000 x 16 L2 vectors (M=8), the `hnsw` strategy
  hnsw_l2_1k.npz         queries, ref_labels_ef{E} / ref_dists_ef{E} = HierarchicalNSW::searchKnn(q, 10), nearest first
  slimzero_l2_1k.graph   HierarchicalNSWSlimZero::saveIndex over the same vectors (`hnsw_slimzero`)
  slimzero_l2_1k.npz     queries, ref_labels_ef{E} = HierarchicalNSWSlimZero::searchKnn(q, 10, out) (unordered rows)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EFS = (10, 40, 100)

if __name__ == "__main__":
    n, dim, nq, k, M = 1000, 16, 64, 10, 8
    base, q = make_dataset(n, nq, dim, metric=0, rank=6, seed=11)
    g = os.path.join(HERE, "hnsw_l2_1k.graph")
    rh.ref_hnsw_build(base, g, M=M, ef_construction=60, branching="4", threads=1)
    ix = rh.RefHnsw(g, dim, n)
    out = {"queries": q, "n": n, "dim": dim, "k": k, "base": base}
    for ef in EFS:
        lab, dist, _ = ix.search(q, k, ef, threads=1)
        out[f"ref_labels_ef{ef}"] = lab
        out[f"ref_dists_ef{ef}"] = dist
    np.savez_compressed(os.path.join(HERE, "hnsw_l2_1k.npz"), **out)
    g = os.path.join(HERE, "slimzero_l2_1k.graph")
    rh.ref_slimzero_build(base, g, M=M, ef_construction=60, branching="4", threads=1)
    iz = rh.RefSlimZero(g, dim, n)
    out = {"queries": q, "n": n, "dim": dim, "k": k}
    for ef in EFS:
        lab, _ = iz.search(q, k, ef)
        out[f"ref_labels_ef{ef}"] = lab
    np.savez_compressed(os.path.join(HERE, "slimzero_l2_1k.npz"), **out)
    print("ok", os.path.getsize(os.path.join(HERE, "hnsw_l2_1k.graph")), os.path.getsize(g))
