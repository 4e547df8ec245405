"""GPU parity of the hnsw_slimq path: traverse_slimq.cu through the C ABI against the C oracle
(oracle/hs_oracle_slimq.c, pinned to the reference by tests/test_oracle_slimq.py) and against the
reference's own golden results.

Bars: against the oracle run with the kernel's association (HSO_ORDER_GPU) the per-query
preparation (rotation, query bit planes, delta/vl/k1xsumq, centroid distances), the result ids,
the exact distances and the per-query counters are BIT-EXACT (rows whose estimates or distances
tie bit-for-bit aside).  Against the reference: same result sets on >= 95 % of queries with the
reference's t_const copied in, recall within 0.5 pp (north_star).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh
from test_oracle_slimq import GOLD, HAVE_REFQ, load_gold, same_sets

pytestmark = pytest.mark.gpu


def open_index(graph, base, t_const=None):
    ix = capi.Index(graph, base.shape[1], kind=capi.HS_KIND_SLIMQ, raw_base=base)
    if t_const is not None:
        ix.query_tconst = t_const
    return ix


def check_against_oracle(ix, o, q, k, ef, min_exact=0.995):
    ix.set_ef(ef)
    lab, dist, cnt = ix.search(q, k, counts=True)
    olab, odist, one, onh, onr = o.search(q, k, ef, order=rh.ORDER_GPU, team=32)
    same = np.all(lab == olab, axis=1)
    assert same.mean() >= min_exact, f"only {same.mean():.4f} of rows identical to the oracle (ef={ef}, k={k})"
    assert np.array_equal(dist[same].view(np.uint32), odist[same].view(np.uint32))
    assert np.array_equal(cnt[same, 0], one[same])
    assert np.array_equal(cnt[same, 1], onh[same])
    return lab, dist


@pytest.mark.parametrize("name", GOLD)
def test_preparation_bit_exact(name):
    g, base, q, graph = load_gold(name)
    t = float(g["t_const"])
    ix = open_index(graph, base, t)
    assert ix.query_tconst == t
    o = rh.OracleQ(graph, base, t_const=t)
    rot, planes, scal, q2c = ix.slimq_prepare(q)
    orot, oplanes, oscal, oq2c = o.prep(q)
    assert np.array_equal(rot.view(np.uint32), orot.view(np.uint32))
    assert np.array_equal(planes, oplanes)
    assert np.array_equal(scal.view(np.uint32), oscal.view(np.uint32))
    assert np.array_equal(q2c.view(np.uint32), oq2c.view(np.uint32))
    # and the reference's own numbers, to the tolerance its build allows
    assert np.abs(rot - g["ref_rotated"]).max() <= 2e-6 * np.abs(g["ref_rotated"]).max()
    assert np.mean(planes == g["ref_planes"]) >= 0.999


@pytest.mark.parametrize("name", GOLD)
@pytest.mark.parametrize("ef", [10, 40, 100, 200])
def test_golden_search(name, ef):
    g, base, q, graph = load_gold(name)
    k = int(g["k"])
    t = float(g["t_const"])
    ix = open_index(graph, base, t)
    o = rh.OracleQ(graph, base, t_const=t)
    lab, dist = check_against_oracle(ix, o, q, k, ef)
    if f"ref_labels_ef{ef}" in g:
        ref = g[f"ref_labels_ef{ef}"]
        full = (lab != 0xFFFFFFFF).all(1)
        assert same_sets(lab[full], ref[full]).mean() >= 0.95


def test_default_tconst_is_deterministic_and_close():
    g, base, q, graph = load_gold("slimq_d96")
    a, b = open_index(graph, base), open_index(graph, base)
    assert a.query_tconst == b.query_tconst
    # the reference's random draws land within a few percent of each other
    assert abs(a.query_tconst / float(g["t_const"]) - 1.0) < 0.05
    a.set_ef(40)
    b.set_ef(40)
    la, da = a.search(q, 10)
    lb, db = b.search(q, 10)
    assert np.array_equal(la, lb) and np.array_equal(da.view(np.uint32), db.view(np.uint32))


def test_k_and_ef_edges():
    g, base, q, graph = load_gold("slimq_d96")
    t = float(g["t_const"])
    ix = open_index(graph, base, t)
    o = rh.OracleQ(graph, base, t_const=t)
    check_against_oracle(ix, o, q, 1, 1)            # ef = k = 1
    check_against_oracle(ix, o, q, 10, 3)           # ef < k: rows padded with 0xFFFFFFFF / inf
    check_against_oracle(ix, o, q, 40, 64)          # k > 32: shared-memory top-k
    check_against_oracle(ix, o, q, 100, 300)        # ef > 128: shared-memory pool
    check_against_oracle(ix, o, q[:1], 10, 50)      # single query
    ix.set_ef(3)
    lab, dist = ix.search(q[:8], 10)
    assert np.all((lab == 0xFFFFFFFF) == np.isinf(dist))


def test_errors_are_loud():
    g, base, q, graph = load_gold("slimq_d96")
    with pytest.raises(capi.HsError):               # rerank rows are mandatory (setDataset)
        capi.Index(graph, 96, kind=capi.HS_KIND_SLIMQ)
    with pytest.raises(capi.HsError):               # wrong dim
        capi.Index(graph, 64, kind=capi.HS_KIND_SLIMQ, raw_base=base[:, :64].copy())
    with pytest.raises(capi.HsError):               # a slimq file is not a slim file
        capi.Index(graph, 96)
    ix = open_index(graph, base)
    with pytest.raises(capi.HsError):
        ix.query_tconst = -1.0


@pytest.mark.skipif(not HAVE_REFQ, reason="needs the compiled reference (oracle/_ref/libhsref_slimq_v4.so)")
def test_larger_index_vs_oracle_and_reference(tmp_path):
    n, nq, dim, k = 30000, 500, 96, 10
    base, q = make_dataset(n, nq, dim, rank=10, seed=3)
    cent, cid = rh.kmeans(base, 16)
    graph = str(tmp_path / "q.graph")
    rh.ref_slimq_build(base, cent, cid, graph, M=16, ef_construction=100, threads=1)
    r = rh.RefSlimQ(graph, base)
    ix = open_index(graph, base, r.t_const)
    o = rh.OracleQ(graph, base, t_const=r.t_const)
    gt, _ = capi.bruteforce_knn(base, q, k)
    for ef in (50, 100, 128, 250):
        lab, dist = check_against_oracle(ix, o, q, k, ef)
        rlab, _ = r.search(q, k, ef)
        rec_g = np.mean([len(set(a) & set(b)) / k for a, b in zip(lab, gt)])
        rec_r = np.mean([len(set(a) & set(b)) / k for a, b in zip(rlab, gt)])
        assert abs(rec_g - rec_r) <= 0.005, (ef, rec_g, rec_r)
        assert same_sets(lab, rlab).mean() >= 0.97
        # returned distances are the exact fp32 distances of the returned rows (1e-5 relative)
        rows = np.searchsorted(np.arange(n), lab)          # labels == internal ids (threads=1 build)
        exact = ((base[rows] - q[:, None, :]) ** 2).sum(-1)
        np.testing.assert_allclose(dist, exact, rtol=1e-5)
    st = ix.stats()
    assert st["n_rerank"] > 0 and st["n_dist"] > st["n_hops"] >= st["n_rerank"]


def test_full_size_properties():
    """Size-independent properties on a 200k x 96 engine-built index (DEEP / MSTuring shape):
    sortedness, distances are the exact distances of the returned rows, recall monotone in ef and
    >= 0.95, determinism under query permutation, and bit-exactness against the oracle on a slice."""
    import tempfile
    n, nq, dim, k = 200000, 2000, 96, 10
    base, q = make_dataset(n, nq, dim, rank=14)
    with tempfile.TemporaryDirectory() as td:
        graph = os.path.join(td, "q.graph")
        capi.build_slimq_graph(base, graph, M=16, ef_construction=100)
        ix = open_index(graph, base)
        gt, _ = capi.bruteforce_knn(base, q[:500], k)
        prev = 0.0
        for ef in (20, 50, 100, 200):
            ix.set_ef(ef)
            lab, dist, cnt = ix.search(q, k, counts=True)
            assert (np.diff(dist, axis=1) >= 0).all() and (lab < n).all()
            assert all(len(set(r)) == k for r in lab[:200])
            rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(lab[:500], gt)])
            assert rec >= prev - 0.002
            prev = rec
            i = np.arange(0, nq, 37)
            true = ((base[lab[i, 0]] - q[i]) ** 2).sum(1)
            np.testing.assert_allclose(dist[i, 0], true, rtol=1e-5)
            assert (cnt[:, 0] > cnt[:, 1]).all() and (cnt[:, 1] >= 1).all()
        assert prev >= 0.95, prev
        a, _ = ix.search(q, k)
        b, _ = ix.search(q[::-1].copy(), k)
        assert np.array_equal(a, b[::-1])
        o = rh.OracleQ(graph, base, t_const=ix.query_tconst)
        check_against_oracle(ix, o, q[:300], k, 100)


def test_threshold_level(tmp_path):
    """threshold_level = 1 index from the engine's builder: bit-exact against the oracle."""
    n, nq, dim, k = 20000, 200, 96, 10
    base, q = make_dataset(n, nq, dim, rank=10, seed=4)
    graph = str(tmp_path / "q.graph")
    capi.build_slimq_graph(base, graph, M=16, ef_construction=100, threshold_level=1)
    ix = open_index(graph, base)
    assert ix.info()["threshold_level"] == 1
    o = rh.OracleQ(graph, base, t_const=ix.query_tconst)
    for ef in (40, 120):
        check_against_oracle(ix, o, q, k, ef)
