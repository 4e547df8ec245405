"""GPU suite: the device-side index builder (hs_build_slim_index_gpu = HNSW build + convertFromHNSW,
hnsw.h:1248-1376 + slim.h:867-1108, csrc/graph_gpu.cu) against the host builder and the reference."""
import os

import numpy as np
import pytest

from conftest import HAVE_REF
from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh

pytestmark = pytest.mark.gpu


def _recall(lab, gt, k):
    return float(np.mean([len(set(a) & set(b)) / k for a, b in zip(lab, gt)]))


def _integrity(path, dim, n, maxM0, maxM):
    """The reference's checkIntegrity (slim.h:2387-2433): ids in range, no self loops, no duplicates."""
    g = capi.HostGraph(path, dim)
    info = g.info()
    assert info["n"] == n
    deg0 = []
    for i in range(0, n, max(1, n // 3000)):
        lvl, _, _ = g.node(i)
        for l in range(lvl + 1):
            row = g.row(i, l)
            assert len(row) <= (maxM0 if l == 0 else maxM)
            assert (row < n).all() and (row != i).all() and len(set(row.tolist())) == len(row)
            if l == 0:
                deg0.append(len(row))
    return info, float(np.mean(deg0))


@pytest.mark.parametrize("n,dim,metric,M,rank", [(30000, 128, 0, 16, 12), (30000, 96, 0, 16, 12), (20000, 200, 0, 16, 12),
                                                 (20000, 64, 1, 16, 10), (15000, 128, 0, 32, 12)])
def test_gpu_built_index_matches_host_built_quality(n, dim, metric, M, rank, tmp_path):
    nq, k = 1000, 10
    base, q = make_dataset(n, nq, dim, metric=metric, rank=rank, seed=5)
    gt, _ = capi.bruteforce_knn(base, q, k, metric=metric)
    host_path = str(tmp_path / "host.graph")
    capi.build_slim_graph(base, host_path, metric=metric, M=M, ef_construction=200)
    host = capi.Index(host_path, dim, metric=metric)
    gpu = capi.Index.build_gpu(base, metric=metric, M=M, ef_construction=200)
    gi, hi = gpu.info(), host.info()
    assert gi["n"] == n and gi["maxlevel"] == hi["maxlevel"] and gi["n_upper"] == hi["n_upper"]   # same level draw
    assert gi["max_deg0"] <= 2 * M
    # the two graphs come from different insertion schedules: equal recall and cost, not equal edges
    assert abs(gi["sum_deg0"] / n - hi["sum_deg0"] / n) < 1.5, (gi["sum_deg0"] / n, hi["sum_deg0"] / n)
    for ef in (20, 50, 100):
        rec, evals = {}, {}
        for name, ix in (("gpu", gpu), ("host", host)):
            ix.set_ef(ef)
            ix.reset_stats()
            lab, _ = ix.search(q, k)
            rec[name] = _recall(lab, gt, k)
            evals[name] = ix.stats()["n_dist"] / nq
        assert rec["gpu"] >= rec["host"] - 0.015, (ef, rec, evals)
        assert evals["gpu"] <= evals["host"] * 1.15, (ef, rec, evals)
    # save -> the file is a valid hnsw_slim .graph: loader round trip gives the same answers
    gpu_path = str(tmp_path / "gpu.graph")
    gpu.save(gpu_path)
    _integrity(gpu_path, dim, n, 2 * M, M)
    again = capi.Index(gpu_path, dim, metric=metric)
    gpu.set_ef(50)
    again.set_ef(50)
    l1, d1 = gpu.search(q, k)
    l2, d2 = again.search(q, k)
    assert np.array_equal(l1, l2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref")
def test_reference_loads_and_searches_a_gpu_built_index(tmp_path):
    """The drop-in direction: HierarchicalNSWSlim::loadIndex + searchKnn (slim.h:753-815, 2030-2131) on a file
    written by hs_save_index return the same neighbours as the engine on the same graph."""
    n, nq, dim, k, ef = 40000, 500, 128, 10, 80
    base, q = make_dataset(n, nq, dim, rank=12, seed=9)
    labels = np.arange(1000, 1000 + n, dtype=np.uint64)            # free-form labels survive the round trip
    gpu = capi.Index.build_gpu(base, M=16, ef_construction=128, labels=labels)
    path = str(tmp_path / "g.graph")
    gpu.save(path)
    gpu.set_ef(ef)
    lab, dist = gpu.search(q, k)
    ref = rh.RefSlim(path, dim, n, 0)
    rlab, _, _ = ref.search(q, k, ef, 1)
    same = np.array([set(a) == set(b) for a, b in zip(lab.tolist(), rlab.tolist())])
    assert same.mean() >= 0.99, same.mean()
    assert lab.min() >= 1000
    gt, _ = capi.bruteforce_knn(base, q, k)
    assert _recall(lab - 1000, gt, k) >= 0.95


def test_gpu_builder_rows_already_on_the_device_and_argument_errors():
    import torch
    n, dim = 12000, 96
    base, q = make_dataset(n, 200, dim, rank=10, seed=3)
    d_base = torch.from_numpy(base).cuda()
    ix = capi.Index.build_gpu(None, base_ptr=d_base.data_ptr(), n=n, dim=dim, M=16, ef_construction=100)
    ix.set_ef(64)
    lab, _ = ix.search(q, 10)
    gt, _ = capi.bruteforce_knn(base, q, 10)
    assert _recall(lab, gt, 10) >= 0.95
    with pytest.raises(capi.HsError):
        capi.Index.build_gpu(base, M=64)                   # beyond the device builder's list capacity
    with pytest.raises(capi.HsError):
        capi.Index.build_gpu(base, M=16, ef_construction=400)


@pytest.mark.parametrize("dim", [96, 128])
def test_gpu_built_slimq_index_matches_host_built_quality(dim, tmp_path):
    """hs_build_slimq_index_gpu (graph + cluster ids + rotation + 1-bit codes + factors, all on the device)
    against hs_build_slimq_graph with the SAME centroids and seed: the per-node payload is a function of
    (row, centroids, rotator bits), so the two indices differ only in their graphs — recall and estimate
    counts must agree like the hnsw_slim builders do."""
    n, nq, k = 30000, 1000, 10
    base, q = make_dataset(n, nq, dim, rank=12, seed=6)
    gt, _ = capi.bruteforce_knn(base, q, k)
    rng = np.random.default_rng(1)
    centroids = base[rng.choice(n, 16, replace=False)].copy()
    cluster_ids = np.argmin(((base[:, None, :] - centroids[None]) ** 2).sum(-1), axis=1).astype(np.uint32)
    host_path = str(tmp_path / "hq.graph")
    capi.build_slimq_graph(base, host_path, M=16, ef_construction=200, centroids=centroids, cluster_ids=cluster_ids)
    host = capi.Index(host_path, dim, kind=capi.HS_KIND_SLIMQ, raw_base=base)
    gpu = capi.Index.build_gpu(base, kind=capi.HS_KIND_SLIMQ, M=16, ef_construction=200, centroids=centroids)
    gi = gpu.info()
    assert gi["kind"] == capi.HS_KIND_SLIMQ and gi["padded_dim_q"] == 128 and gi["num_cluster"] == 16
    for ef in (30, 60, 120):
        rec, est = {}, {}
        for name, ix in (("gpu", gpu), ("host", host)):
            ix.set_ef(ef)
            ix.reset_stats()
            lab, dist = ix.search(q, k)
            rec[name] = _recall(lab, gt, k)
            est[name] = ix.stats()["n_dist"] / nq
            assert (np.diff(dist, axis=1) >= 0).all()
        assert rec["gpu"] >= rec["host"] - 0.015, (ef, rec, est)
        assert est["gpu"] <= est["host"] * 1.15, (ef, rec, est)
    # centroids picked by the builder itself (k-means over a host-side sample)
    auto = capi.Index.build_gpu(base, kind=capi.HS_KIND_SLIMQ, M=16, ef_construction=200)
    auto.set_ef(60)
    host.set_ef(60)
    la, _ = auto.search(q, k)
    lh, _ = host.search(q, k)
    assert _recall(la, gt, k) >= _recall(lh, gt, k) - 0.02
