"""GPU: the tcgen05 exact-kNN path (bruteforce_tc.cu) returns bit for bit what the fp32 scan
kernel (bruteforce.cu, itself bit-exact against the oracle / the reference's BruteforceSearch,
tests/test_gpu_exact.py) returns: ids AND distances, for every shape, metric and k."""
import os

import numpy as np
import pytest

from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh

pytestmark = pytest.mark.gpu


def run(base, q, k, metric, tc):
    os.environ["HS_BF_TC"] = "1" if tc else "0"
    os.environ["HS_BF_TC_STATS"] = "1"
    try:
        lab, dist = capi.bruteforce_knn(base, q, k, metric=metric)
        return lab, dist, capi.bf_tc_fallback()
    finally:
        del os.environ["HS_BF_TC"]
        del os.environ["HS_BF_TC_STATS"]


@pytest.mark.parametrize("n,nq,dim,metric,k", [
    (20000, 300, 128, 0, 10),      # SIFT shape, nq not a multiple of 128
    (20000, 300, 128, 0, 100),     # the ground-truth k
    (13001, 129, 96, 0, 10),       # DEEP shape, ragged n and nq
    (9000, 256, 32, 0, 10),        # a single k-block
    (9000, 130, 200, 0, 10),       # dim not a multiple of 32 (padded k-block)
    (6000, 128, 960, 0, 10),       # GIST shape: 30 k-blocks
    (20000, 300, 768, 1, 10),      # COHERE shape, inner product on unit vectors
    (20000, 64, 128, 1, 100),
])
def test_tc_equals_scan(n, nq, dim, metric, k):
    base, q = make_dataset(n, nq, dim, metric=metric, rank=12, seed=11)
    lab0, dist0, fb0 = run(base, q, k, metric, tc=False)
    lab1, dist1, fb1 = run(base, q, k, metric, tc=True)
    assert fb0 == -1 and fb1 >= 0
    assert np.array_equal(lab0, lab1)
    assert np.array_equal(dist0.view(np.uint32), dist1.view(np.uint32))
    assert fb1 <= nq // 10, f"{fb1} of {nq} queries fell back to the scan kernel"


def test_tc_against_oracle():
    base, q = make_dataset(30000, 200, 128, rank=12, seed=5)
    lab, dist, fb = run(base, q, 10, 0, tc=True)
    ol, od = rh.oracle_bruteforce(base, q, 10, order=rh.ORDER_SEQFMA)
    assert np.array_equal(lab, ol[:, ::-1])
    assert np.array_equal(dist.view(np.uint32), od[:, ::-1].copy().view(np.uint32))


def test_tc_ties_and_duplicates_use_the_fallback_correctly():
    """Many exactly equal rows: the candidate heaps cannot prove sufficiency, the affected queries
    go through the scan kernel, results stay identical (ties -> smaller label)."""
    rng = np.random.default_rng(3)
    proto = rng.standard_normal((50, 64)).astype(np.float32)
    base = np.repeat(proto, 400, axis=0)                      # 20000 rows, 400 copies of each
    base = base[rng.permutation(base.shape[0])]
    q = proto[:40] + 0.01 * rng.standard_normal((40, 64)).astype(np.float32)
    q = np.concatenate([q, rng.standard_normal((100, 64)).astype(np.float32)])
    lab0, dist0, _ = run(base, q, 10, 0, tc=False)
    lab1, dist1, fb = run(base, q, 10, 0, tc=True)
    assert fb > 0
    assert np.array_equal(lab0, lab1)
    assert np.array_equal(dist0.view(np.uint32), dist1.view(np.uint32))


def test_tc_full_size_properties():
    """1M x 128 (the ground-truth job of the headline config), 1024 queries: sorted rows, exact
    distances of the returned rows, and agreement with the scan kernel on a slice."""
    base, q = make_dataset(1_000_000, 1024, 128, rank=14)
    lab1, dist1, fb = run(base, q, 100, 0, tc=True)
    assert fb <= 10
    assert (np.diff(dist1, axis=1) >= 0).all() and (lab1 < base.shape[0]).all()
    i = np.arange(0, 1024, 41)
    true = ((base[lab1[i, 0]] - q[i]) ** 2).sum(1)
    np.testing.assert_allclose(dist1[i, 0], true, rtol=1e-5)
    lab0, dist0, _ = run(base, q[:128], 100, 0, tc=False)
    assert np.array_equal(lab0, lab1[:128]) and np.array_equal(dist0.view(np.uint32), dist1[:128].view(np.uint32))
