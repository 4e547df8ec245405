"""The two strategies next to hnsw_slim that share its search code:

* `hnsw`          — the un-pruned hnswlib index (hnsw_strategy.h:15-61).  HierarchicalNSW::searchKnn
                    (hnsw.h:1378-1440) is the slim search with threshold_level 0 on full lists; the
                    engine loads the upstream file format with kind = HS_KIND_HNSW.
* `hnsw_slimzero` — HierarchicalNSWSlimZero (hnswalg_slimzero.h): a different pruning at build time,
                    the same file format (:701-735) and the same searchKnn (:1675-1771) as hnsw_slim;
                    the engine loads it as HS_KIND_SLIM.

CPU part: the oracle and the loader against golden vectors from the reference and against the live
reference.  GPU part (-m gpu): the CUDA traversal against the oracle (bit-exact) and the reference.
"""
import os

import numpy as np
import pytest

from conftest import CACHE, GOLDEN, needs_ref_hnsw
from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh

EFS = (10, 40, 100)


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz")), os.path.join(GOLDEN, name + ".graph")


# ------------------------------------------------------------------ CPU: oracle + loader --------
@pytest.mark.parametrize("order", [rh.ORDER_SEQ, rh.ORDER_REF, rh.ORDER_GPU])
def test_golden_hnsw_oracle_matches_reference(order):
    z, graph = _golden("hnsw_l2_1k")
    orc = rh.Oracle(graph, int(z["dim"]), 0, hnsw=True)
    for ef in EFS:
        lab, dist, _, _ = orc.search(z["queries"], int(z["k"]), ef, order=order, team=8, threads=1)
        ref_l, ref_d = z[f"ref_labels_ef{ef}"], z[f"ref_dists_ef{ef}"]
        same = np.array([set(a) == set(b) for a, b in zip(lab, ref_l)])
        assert same.mean() >= 0.97, (ef, same.mean())
        np.testing.assert_allclose(dist[same], ref_d[same], rtol=1e-5)       # both nearest first


def test_golden_slimzero_oracle_matches_reference():
    z, graph = _golden("slimzero_l2_1k")
    orc = rh.Oracle(graph, int(z["dim"]), 0)                                   # same format as hnsw_slim
    for ef in EFS:
        lab, _, _, _ = orc.search(z["queries"], int(z["k"]), ef, order=rh.ORDER_REF, threads=1)
        same = np.array([set(a) == set(b) for a, b in zip(lab, z[f"ref_labels_ef{ef}"])])
        assert same.mean() >= 0.97, (ef, same.mean())


def test_hnsw_loader_matches_oracle_lists():
    """hs_debug_flatten(kind = HS_KIND_HNSW): every level list, level, label and vector of the
    upstream-format file, as the oracle's own parser reads them."""
    z, graph = _golden("hnsw_l2_1k")
    dim, n = int(z["dim"]), int(z["n"])
    g = capi.HostGraph(graph, dim, kind=capi.HS_KIND_HNSW)
    orc = rh.Oracle(graph, dim, 0, hnsw=True)
    info, oinfo = g.info(), orc.info()
    assert (info["n"], info["maxlevel"], info["enterpoint"], info["maxM"], info["maxM0"], info["threshold_level"]) == \
        (oinfo["n"], oinfo["maxlevel"], oinfo["enterpoint"], oinfo["maxM"], oinfo["maxM0"], 0)
    assert info["kind"] == capi.HS_KIND_HNSW
    for i in range(n):
        level, label, _ = orc.node(i, 0)
        glevel, glabel, gvec = g.node(i)
        assert (glevel, glabel) == (level, label)
        assert np.array_equal(gvec[:dim], z["base"][label]) and not gvec[dim:].any()
        for l in range(level + 1):
            assert np.array_equal(g.row(i, l), orc.node(i, l)[2]), (i, l)


def test_hnsw_loader_rejects_wrong_kind_and_truncation(tmp_path):
    _, graph = _golden("hnsw_l2_1k")
    with pytest.raises(capi.HsError):
        capi.HostGraph(graph, 16, kind=capi.HS_KIND_SLIM)       # an hnsw file is not a CHAL file
    with pytest.raises(capi.HsError):
        capi.HostGraph(graph, 24, kind=capi.HS_KIND_HNSW)       # wrong dim
    img = open(graph, "rb").read()
    cut = tmp_path / "cut.graph"
    cut.write_bytes(img[: len(img) - 9])
    with pytest.raises(capi.HsError):
        capi.HostGraph(str(cut), 16, kind=capi.HS_KIND_HNSW)


class _Corpus:
    pass


_cache = {}


def _corpus(kind: str, n=20000, nq=300, dim=32, metric=0):
    key = (kind, n, nq, dim, metric)
    if key not in _cache:
        c = _Corpus()
        c.n, c.dim, c.metric = n, dim, metric
        c.base, c.queries = make_dataset(n, nq, dim, metric=metric, rank=8, seed=3)
        os.makedirs(CACHE, exist_ok=True)
        c.graph = os.path.join(CACHE, f"{kind}_{n}_{dim}_{metric}.graph")
        if not os.path.exists(c.graph):
            tmp = c.graph + f".tmp{os.getpid()}"
            if kind == "hnsw":
                rh.ref_hnsw_build(c.base, tmp, metric=metric, M=16, ef_construction=200)
            else:
                rh.ref_slimzero_build(c.base, tmp, metric=metric, M=16, ef_construction=200)
            os.replace(tmp, c.graph)
        _cache[key] = c
    return _cache[key]


@needs_ref_hnsw
@pytest.mark.parametrize("metric", [0, 1])
def test_live_hnsw_oracle_matches_reference(metric):
    c = _corpus("hnsw", metric=metric)
    ref = rh.RefHnsw(c.graph, c.dim, c.n, metric)
    orc = rh.Oracle(c.graph, c.dim, metric, hnsw=True)
    for ef in (20, 100):
        rl, rd, _ = ref.search(c.queries, 10, ef)
        ol, od, _, _ = orc.search(c.queries, 10, ef, order=rh.ORDER_REF)
        same = np.array([set(a) == set(b) for a, b in zip(ol, rl)])
        assert same.mean() >= 0.99, (ef, same.mean())
        np.testing.assert_allclose(od[same], rd[same], rtol=1e-5, atol=1e-6)


@needs_ref_hnsw
def test_live_slimzero_oracle_matches_reference():
    c = _corpus("slimzero")
    ref = rh.RefSlimZero(c.graph, c.dim, c.n)
    orc = rh.Oracle(c.graph, c.dim, 0)
    for ef in (20, 100):
        rl, _ = ref.search(c.queries, 10, ef)
        ol, _, _, _ = orc.search(c.queries, 10, ef, order=rh.ORDER_REF)
        same = np.array([set(a) == set(b) for a, b in zip(ol, rl)])
        assert same.mean() >= 0.99, (ef, same.mean())


# ------------------------------------------------------------------ GPU: the CUDA traversal -----
def _check_gpu_vs_oracle(graph, dim, metric, kind, queries, k, ef, hnsw):
    ix = capi.Index(graph, dim, kind=kind, metric=metric)
    ix.set_ef(ef)
    lab, dist, cnt = ix.search(queries, k, counts=True)
    ol, od, ond, onh = rh.Oracle(graph, dim, metric, hnsw=hnsw).search(queries, k, ef, order=rh.ORDER_GPU, team=8)
    same = np.all(lab == ol, axis=1)
    assert same.mean() >= 0.999, same.mean()
    assert np.array_equal(dist[same].view(np.uint32), od[same].view(np.uint32))
    assert (cnt[same, 0] == ond[same]).mean() >= 0.999 and (cnt[same, 1] == onh[same]).mean() >= 0.999
    return lab, dist


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,hnsw", [("hnsw_l2_1k", capi.HS_KIND_HNSW, True),
                                            ("slimzero_l2_1k", capi.HS_KIND_SLIM, False)])
def test_gpu_golden_graphs_match_oracle_and_reference(name, kind, hnsw):
    z, graph = _golden(name)
    for ef in EFS:
        lab, _ = _check_gpu_vs_oracle(graph, int(z["dim"]), 0, kind, z["queries"], int(z["k"]), ef, hnsw)
        same = np.array([set(a) == set(b) for a, b in zip(lab, z[f"ref_labels_ef{ef}"])])
        assert same.mean() >= 0.97, (ef, same.mean())


@pytest.mark.gpu
@needs_ref_hnsw
@pytest.mark.parametrize("metric", [0, 1])
def test_gpu_hnsw_strategy_matches_oracle_and_reference(metric):
    """`--solve_strategy=hnsw` on the GPU: bit-exact vs the oracle, same results as the reference's
    HierarchicalNSW::searchKnn (recall within 0.5 pp, distances within 1e-5 relative)."""
    c = _corpus("hnsw", metric=metric)
    for ef in (10, 100, 200, 300):
        lab, dist = _check_gpu_vs_oracle(c.graph, c.dim, metric, capi.HS_KIND_HNSW, c.queries, 10, ef, True)
    rl, rd, _ = rh.RefHnsw(c.graph, c.dim, c.n, metric).search(c.queries, 10, 300)
    same = np.array([set(a) == set(b) for a, b in zip(lab, rl)])
    assert same.mean() >= 0.99, same.mean()
    np.testing.assert_allclose(dist[same], rd[same], rtol=1e-5, atol=1e-6)
    gt, _ = capi.bruteforce_knn(c.base, c.queries, 10, metric=metric)
    rec = lambda l: np.mean([len(set(a) & set(b)) / 10 for a, b in zip(l, gt)])
    assert abs(rec(lab) - rec(rl)) <= 0.005


@pytest.mark.gpu
@needs_ref_hnsw
def test_gpu_slimzero_strategy_matches_oracle_and_reference():
    c = _corpus("slimzero")
    for ef in (10, 100, 300):
        lab, _ = _check_gpu_vs_oracle(c.graph, c.dim, 0, capi.HS_KIND_SLIM, c.queries, 10, ef, False)
    rl, _ = rh.RefSlimZero(c.graph, c.dim, c.n).search(c.queries, 10, 300)
    same = np.array([set(a) == set(b) for a, b in zip(lab, rl)])
    assert same.mean() >= 0.99, same.mean()
