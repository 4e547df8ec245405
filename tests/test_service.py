"""hs_service: the serving front (single queries from many threads -> batches -> hs_search_batch), the
GPU-engine form of the reference server's /query, /setEf and update handlers (hnsw_slim_server.cc:69-142,
hnsw_slim_server_patch.cc:133-228).

CPU part: the batching logic over a stand-in backend (exact search in numpy) — every caller gets ITS answer,
batches form under load, k is not mixed inside a batch, errors propagate.  GPU part: answers equal the batched
call's, set_ef and patches land between batches."""
import os
import threading
import time

import numpy as np
import pytest

from conftest import GOLDEN
from hnsw_slim_b200 import capi
from oracle import refharness as rh


class ExactBackend:
    def __init__(self, base, delay=0.0):
        self.base, self.delay = base, delay
        self.calls = []                       # (nq, k) per launch
        self.fail = False

    def __call__(self, q, k):
        self.calls.append((q.shape[0], k))
        if self.fail:
            raise RuntimeError("boom")
        if self.delay:
            time.sleep(self.delay)
        d = ((q[:, None, :] - self.base[None, :, :]) ** 2).sum(-1)
        lab = np.argsort(d, axis=1, kind="stable")[:, :k].astype(np.uint32)
        return lab, np.take_along_axis(d, lab.astype(np.int64), axis=1).astype(np.float32)


def _hammer(svc, queries, ks, n_threads):
    out = [None] * len(queries)
    err = []

    def work(t):
        try:
            for i in range(t, len(queries), n_threads):
                out[i] = svc.query(queries[i], ks[i], want_dists=True)
        except Exception as e:                # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not any(t.is_alive() for t in th), "a request never returned"
    return out, err


def test_service_every_caller_gets_its_own_answer():
    rng = np.random.default_rng(0)
    base = rng.standard_normal((500, 8)).astype(np.float32)
    q = rng.standard_normal((400, 8)).astype(np.float32)
    be = ExactBackend(base, delay=0.002)
    svc = capi.Service(be, max_batch=64, k_max=10, dim=8)
    ks = [5] * len(q)
    out, err = _hammer(svc, q, ks, 16)
    assert not err
    want_l, want_d = ExactBackend(base)(q, 5)
    for i, (lab, dist) in enumerate(out):
        assert np.array_equal(lab, want_l[i]) and np.allclose(dist, want_d[i])
    st = svc.stats()
    assert st["queries"] == 400 and st["batches"] == len(be.calls)
    assert st["batches"] < 400 and st["max_batch"] > 1            # requests coalesced while a batch was running
    assert st["max_batch"] <= 64 and all(n <= 64 for n, _ in be.calls)
    svc.close()


def test_service_does_not_mix_k_inside_a_batch():
    rng = np.random.default_rng(1)
    base = rng.standard_normal((300, 4)).astype(np.float32)
    q = rng.standard_normal((240, 4)).astype(np.float32)
    be = ExactBackend(base, delay=0.001)
    svc = capi.Service(be, max_batch=32, k_max=20, dim=4)
    ks = [(3, 7, 20)[i % 3] for i in range(len(q))]
    out, err = _hammer(svc, q, ks, 12)
    assert not err
    for i, (lab, dist) in enumerate(out):
        wl, _ = ExactBackend(base)(q[i:i + 1], ks[i])
        assert lab.shape == (ks[i],) and np.array_equal(lab, wl[0])
    assert sum(n for n, _ in be.calls) == len(q)
    assert {k for _, k in be.calls} == {3, 7, 20}
    with pytest.raises(capi.HsError):
        svc.query(q[0], 21)                                      # above k_max
    svc.close()


def test_service_max_wait_collects_a_batch():
    rng = np.random.default_rng(2)
    base = rng.standard_normal((100, 4)).astype(np.float32)
    q = rng.standard_normal((40, 4)).astype(np.float32)
    be = ExactBackend(base)
    svc = capi.Service(be, max_batch=8, max_wait_us=200000, k_max=4, dim=4)
    t0 = time.time()
    out, err = _hammer(svc, q[:8], [4] * 8, 8)                    # 8 concurrent requests fill the batch: no waiting
    assert not err and time.time() - t0 < 0.19
    assert be.calls == [(8, 4)]
    t0 = time.time()
    svc.query(q[9], 4)                                           # a lone request waits out the window
    assert time.time() - t0 >= 0.15
    svc.close()


def test_service_backend_failure_reaches_the_callers():
    rng = np.random.default_rng(3)
    base = rng.standard_normal((50, 4)).astype(np.float32)
    be = ExactBackend(base)
    svc = capi.Service(be, max_batch=4, k_max=4, dim=4)
    assert svc.query(base[3], 1)[0] == 3
    be.fail = True
    with pytest.raises(capi.HsError):
        svc.query(base[4], 1)
    be.fail = False
    assert svc.query(base[5], 1)[0] == 5                         # and the service carries on
    with pytest.raises(capi.HsError):
        svc.patch(b"\0" * 24)                                    # no index behind a stand-in backend
    svc.close()


def test_service_argument_errors():
    import ctypes as C
    L, out = capi.lib(), C.c_void_p()
    assert L.hs_service_create(None, 16, 0, 10, C.byref(out)) == -1
    assert L.hs_service_query(None, None, 1, None, None) == -1
    assert L.hs_service_set_ef(None, 10) == -1
    assert L.hs_service_get_stats(None, None) == -1
    L.hs_service_free(None)


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_service_answers_like_the_batched_call():
    from conftest import get_corpus
    c = get_corpus(n=20000, nq=600, dim=128)
    ix = capi.Index(c.graph, c.dim)
    ix.set_ef(64)
    want_l, want_d = ix.search(c.queries, 10)
    svc = capi.Service(ix, max_batch=256, k_max=10)
    out, err = _hammer(svc, c.queries, [10] * len(c.queries), 32)
    assert not err
    for i, (lab, dist) in enumerate(out):
        assert np.array_equal(lab, want_l[i]) and np.array_equal(dist.view(np.uint32), want_d[i].view(np.uint32))
    st = svc.stats()
    assert st["queries"] == 600 and st["batches"] < 600
    svc.set_ef(16)                                               # setEf for the batches launched from now on
    ix2 = capi.Index(c.graph, c.dim)
    ix2.set_ef(16)
    w2, _ = ix2.search(c.queries[:50], 10)
    for i in range(50):
        assert np.array_equal(svc.query(c.queries[i], 10), w2[i])
    svc.close()


@pytest.mark.gpu
def test_gpu_service_patch_lands_between_batches():
    z = np.load(os.path.join(GOLDEN, "patch_l2_1k.npz"))
    graph = os.path.join(GOLDEN, "patch_l2_1k.graph")
    streams = [z["patch0"].tobytes(), z["patch1"].tobytes()]
    q, dim = z["queries"], int(z["dim"])
    ix = capi.Index.load_reserve(graph, dim, 1200)
    ix.set_ef(40)
    svc = capi.Service(ix, max_batch=32, k_max=10)
    stages = []
    for s in range(3):
        orc = rh.Oracle(graph, dim)
        for j in range(s):
            orc.patch(streams[j], rows=None if j == 1 else z["base"], inline=(j == 1))
        stages.append(orc.search(q, 10, 40, order=rh.ORDER_GPU, team=8)[0])
    stop = threading.Event()
    bad = []

    def client(t):
        i = t
        while not stop.is_set():
            lab = svc.query(q[i % len(q)], 10)
            if not any(np.array_equal(lab, st[i % len(q)]) for st in stages):
                bad.append(i)                                     # an answer from a half-patched index
            i += 4

    th = [threading.Thread(target=client, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    time.sleep(0.05)
    info0 = svc.patch(streams[0], rows=z["base"])
    time.sleep(0.05)
    info1 = svc.patch(streams[1], inline=True)
    time.sleep(0.05)
    stop.set()
    [t.join(timeout=60) for t in th]
    assert not bad and (info0["n_after"], info1["n_after"]) == (1050, 1200)
    for i in range(len(q)):
        assert np.array_equal(svc.query(q[i], 10), stages[2][i])
    assert svc.stats()["patches"] == 2
    svc.close()
