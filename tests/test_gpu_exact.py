"""GPU parity for the exact paths: brute-force kNN (the recall_knn / ground-truth path), the
cross-shard top-k merge and recall — all through the C ABI, against the oracle and the reference.

Bar (north_star): exact-kNN ids bit-exact with the reference's CPU ground truth except for
distance ties.  Against the oracle run with the kernel's own fp32 association (one fma chain in
index order, HSO_ORDER_SEQFMA) ids AND distances are bit-exact with no exception."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, get_corpus, needs_ref
from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nq,dim,k,metric", [
    (5000, 130, 128, 10, 0), (5000, 33, 96, 100, 0), (3000, 7, 960, 10, 0), (4000, 64, 768, 10, 1),
    (300, 5, 17, 300, 0), (64, 1, 4, 1, 0), (10000, 257, 100, 50, 0)])
def test_bruteforce_bit_exact_vs_oracle(n, nq, dim, k, metric):
    base, q = make_dataset(n, nq, dim, metric=metric, rank=min(8, dim))
    lab, dist = capi.bruteforce_knn(base, q, k, metric=metric)
    olab, odist = rh.oracle_bruteforce(base, q, k, metric=metric, order=rh.ORDER_SEQFMA)   # farthest first
    assert np.array_equal(lab, olab[:, ::-1])
    assert np.array_equal(dist.view(np.uint32), odist[:, ::-1].view(np.uint32))
    assert (np.diff(dist, axis=1) >= 0).all()


def test_bruteforce_tie_rule_duplicates():
    """Equal distances: smaller label wins at the k-th boundary (bruteforce.h:109 pair order)."""
    base = np.tile(np.arange(8, dtype=np.float32)[:, None], (50, 4))          # every row 50 times
    q = np.zeros((3, 4), np.float32)
    lab, dist = capi.bruteforce_knn(base, q, 60)
    olab, odist = rh.oracle_bruteforce(base, q, 60, order=rh.ORDER_SEQFMA)
    assert np.array_equal(lab, olab[:, ::-1])
    assert np.array_equal(lab[0, :50], np.arange(0, 400, 8, dtype=np.uint32))  # the 50 copies of row 0, by label


@needs_ref
def test_bruteforce_vs_reference_ground_truth():
    c = get_corpus(n=20000, nq=300, dim=32)
    lab, dist = capi.bruteforce_knn(c.base, c.queries, 100)
    ref, rdist, _ = rh.ref_bruteforce(c.base, c.queries, 100, want_dists=True)
    ref, rdist = ref[:, ::-1], rdist[:, ::-1]
    diff = lab != ref
    assert diff.mean() < 0.002
    # every disagreement is a distance tie (to fp32 rounding of the two associations)
    for i, j in zip(*np.nonzero(diff)):
        assert abs(dist[i, j] - rdist[i, j]) <= 1e-5 * abs(rdist[i, j])
        assert set(lab[i]) ^ set(ref[i]) == set() or abs(dist[i, -1] - rdist[i, -1]) <= 1e-5 * abs(rdist[i, -1])
    np.testing.assert_allclose(dist, rdist, rtol=1e-5)


@pytest.mark.parametrize("name,metric", [("slim_l2_2k", 0), ("slim_ip_1k", 1)])
def test_golden_reference_vectors(name, metric):
    """Committed outputs of the reference itself: traversal labels per ef and brute-force rows."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    graph = os.path.join(GOLDEN, name + ".graph")
    dim, n, k = int(z["dim"]), int(z["n"]), int(z["k"])
    ix = capi.Index(graph, dim, metric=metric)
    for ef in (10, 40, 100):
        ix.set_ef(ef)
        lab, dist, cnt = ix.search(z["queries"], k, counts=True)
        ref = z[f"ref_labels_ef{ef}"]
        same = np.array([set(a) == set(b) for a, b in zip(lab, ref)])
        assert same.mean() >= 0.97, (ef, same.mean())
        assert (cnt[same, 0] == z[f"ref_counts_ef{ef}"][same]).all()
    orc = rh.Oracle(graph, dim, metric)
    base = np.stack([orc.vector(i) for i in range(n)])
    lab, dist = capi.bruteforce_knn(base, z["queries"], 100, metric=metric)
    ref = z["ref_gt100"][:, ::-1]
    assert (lab == ref).mean() >= 0.995
    assert all(len(set(a) ^ set(b)) <= 2 for a, b in zip(lab, ref))


def test_topk_merge_matches_sorting():
    rng = np.random.default_rng(3)
    for parts, nq, k in [(8, 100, 10), (2, 17, 100), (1, 5, 3), (8, 1000, 10)]:
        d = rng.random((parts, nq, k)).astype(np.float32)
        d[rng.random(d.shape) < 0.05] = 0.25                       # ties across parts
        d = np.sort(d, axis=2)
        lab = rng.permutation(parts * nq * k).astype(np.uint32).reshape(parts, nq, k)
        lab[0, 0, -1] = 0xFFFFFFFF                                  # padding entries are skipped
        d[0, 0, -1] = np.inf
        tl, td = torch.from_numpy(lab.view(np.int32)).cuda(), torch.from_numpy(d).cuda()
        ol = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        capi.topk_merge_device(tl.data_ptr(), td.data_ptr(), parts, nq, k, ol.data_ptr(), od.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got_l, got_d = ol.cpu().numpy().view(np.uint32), od.cpu().numpy()
        for qi in range(nq):
            pairs = sorted((float(dd), int(ll)) for dd, ll in zip(d[:, qi].ravel(), lab[:, qi].ravel())
                           if ll != 0xFFFFFFFF)[:k]
            pairs += [(float("inf"), 0xFFFFFFFF)] * (k - len(pairs))      # fewer than k valid entries: padded
            assert [p[1] for p in pairs] == got_l[qi].tolist()
            assert [np.float32(p[0]) for p in pairs] == got_d[qi].tolist()


def test_recall_matches_oracle():
    base, q = make_dataset(3000, 60, 32, rank=8)
    gt, _ = capi.bruteforce_knn(base, q, 100)
    rng = np.random.default_rng(0)
    knn = gt[:, :10].copy()
    knn[rng.random(knn.shape) < 0.3] = 2999                     # spoil 30 %
    for arr in (gt, gt[:, ::-1].copy(), gt[:, rng.permutation(100)]):
        r = capi.recall(base, q, knn, arr, 10)
        assert abs(r - rh.oracle_recall(base, q, knn, arr, 10)) < 1e-12
    assert capi.recall(base, q, gt[:, :10].copy(), gt, 10) == 1.0
    # ... and against the reference's own SolveStrategy::recall, executed (prints 6 significant digits)
    if rh.ref_slim_path() is not None and hasattr(rh.slim_lib(), "ref_strategy_recall"):
        knn2 = gt[:, :10].copy()
        spoil = rng.random(knn2.shape) < 0.3
        knn2[spoil] = (2900 + np.arange(10)[None, :].repeat(len(knn2), 0))[spoil]      # distinct wrong answers per row
        want = rh.ref_strategy_recall(base, q, knn2, gt, 10)
        assert abs(capi.recall(base, q, knn2, gt, 10) - want) <= 1e-5, want
