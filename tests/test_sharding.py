"""CPU suite, part 3: the multi-GPU host logic under a real 2-process gloo group
(shard assignment, all-gather, merge order).  The GPU path swaps in NCCL and the CUDA merge."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from hnsw_slim_b200 import sharding


def test_shard_ranges_and_assignment():
    assert sharding.shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    r = sharding.shard_ranges(100_000_000, 8)
    assert r[0] == (0, 12_500_000) and r[-1][1] == 100_000_000
    assert sharding.shards_of_rank(8, 1, 2) == [4, 5, 6, 7]
    assert sharding.shards_of_rank(8, 3, 8) == [3]
    with pytest.raises(ValueError):
        sharding.shards_of_rank(8, 0, 3)


def test_merge_numpy_semantics():
    lab = np.array([[[1, 5, 0xFFFFFFFF]], [[2, 9, 7]]], dtype=np.uint32)       # parts=2, nq=1, k=3
    dst = np.array([[[0.5, 0.7, np.inf]], [[0.5, 0.6, 0.9]]], dtype=np.float32)
    l, d = sharding.merge_numpy(lab, dst, 3)
    assert l.tolist() == [[1, 2, 9]] and d.tolist() == [[0.5, 0.5, np.float32(0.6)]]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                      # same stream on every rank
    nq, k, n = 50, 10, 4000
    base = rng.standard_normal((n, 8)).astype(np.float32)
    q = rng.standard_normal((nq, 8)).astype(np.float32)
    lo, hi = sharding.shard_ranges(n, world)[sharding.shards_of_rank(world, rank, world)[0]]
    d = ((base[None, lo:hi] - q[:, None]) ** 2).sum(-1).astype(np.float32)   # this rank's shard, exact
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    loc_l = torch.from_numpy((idx + lo).astype(np.int32))                     # GLOBAL labels
    loc_d = torch.from_numpy(np.take_along_axis(d, idx, 1))

    def merge(all_l, all_d, kk):
        return sharding.merge_numpy(all_l.numpy().view(np.uint32), all_d.numpy(), kk)

    got_l, got_d = sharding.gather_and_merge(loc_l, loc_d, k, merge=merge)
    full = ((base[None] - q[:, None]) ** 2).sum(-1).astype(np.float32)
    want = np.stack([np.lexsort((np.arange(n), full[i]))[:k] for i in range(nq)])
    ok = np.array_equal(got_l, want.astype(np.uint32)) and np.allclose(got_d, np.take_along_axis(full, want, 1))
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_merge_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["slim", "slimq"])
def test_sharded_index_local_shards_on_gpu(kind, tmp_path):
    """ShardedIndex on one GPU (world size 1): four sub-graphs with GLOBAL labels, every shard
    searched on the whole batch (overlapping launches), hs_topk_merge_device over the local results.
    The merged rows equal the (dist, label)-sorted union of the per-shard rows, and the recall is at
    least that of one graph over the whole corpus at the same ef."""
    import torch
    from hnsw_slim_b200 import capi, sharding
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k, ef, S = 24000, 400, 96, 10, 80, 4
    base, q = make_dataset(n, nq, dim, rank=10, seed=21)
    ranges = sharding.shard_ranges(n, S)
    paths, raws = [], []
    for s, (lo, hi) in enumerate(ranges):
        p = str(tmp_path / f"s{s}.graph")
        labels = np.arange(lo, hi, dtype=np.uint64)
        if kind == "slimq":
            capi.build_slimq_graph(base[lo:hi], p, M=16, ef_construction=100, labels=labels)
            raws.append(base[lo:hi])
        else:
            capi.build_slim_graph(base[lo:hi], p, M=16, ef_construction=100, labels=labels)
        paths.append(p)
    K = capi.HS_KIND_SLIMQ if kind == "slimq" else capi.HS_KIND_SLIM
    ix = sharding.ShardedIndex(paths, dim, device=0, kind=K, raw_bases=raws if kind == "slimq" else None)
    ix.set_ef(ef)
    dq = torch.from_numpy(q).cuda()
    lab, dist = ix.search(dq, nq, k)
    ix.join()
    lab, dist = lab.cpu().numpy().view(np.uint32), dist.cpu().numpy()
    # per-shard results through the plain single-index API, merged on the host
    parts_l, parts_d = [], []
    for i, p in enumerate(paths):
        one = capi.Index(p, dim, kind=K, raw_base=raws[i] if kind == "slimq" else None)
        one.set_ef(ef)
        l, d = one.search(q, k)
        parts_l.append(l)
        parts_d.append(d)
    want_l, want_d = sharding.merge_numpy(np.stack(parts_l), np.stack(parts_d), k)
    assert np.array_equal(lab, want_l)
    assert np.array_equal(dist.view(np.uint32), want_d.view(np.uint32))
    assert (np.diff(dist, axis=1) >= 0).all()
    gt, _ = capi.bruteforce_knn(base, q, k)
    rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(lab, gt)])
    assert rec >= 0.97, rec
    # a stream of batches, more than the group's depth, enqueued without waiting: same rows after join()
    outs = [ix.search(dq, nq, k) for _ in range(11)]
    ix.join()
    for pl, pd in outs:
        assert np.array_equal(pl.cpu().numpy().view(np.uint32), want_l)
        assert np.array_equal(pd.cpu().numpy().view(np.uint32), want_d.view(np.uint32))
    # the all-gather form of the exchange (world size 1: no collective) gives the same rows
    nl, nd = ix.search(dq, nq, k, exchange="nccl")
    torch.cuda.synchronize()
    assert np.array_equal(nl.cpu().numpy().view(np.uint32), want_l)
    assert np.array_equal(nd.cpu().numpy().view(np.uint32), want_d.view(np.uint32))
    ix.close()


@pytest.mark.gpu
def test_shard_group_skewed_shards_complete_before_merge(tmp_path):
    """The launches of a batch overlap (programmatic stream serialization) and may FINISH in any order:
    a large first shard followed by tiny ones must still be complete when the merge reads the table —
    the batch is signalled by the last warp to finish over all launches, not by the last launch."""
    import torch
    from hnsw_slim_b200 import capi, sharding
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k = 60000, 2000, 128, 10
    base, q = make_dataset(n, nq, dim, rank=10, seed=77)
    cuts = [0, 56000, 58000, 59000, 60000]               # 56000 rows, then 2000 / 1000 / 1000
    paths = []
    for s in range(4):
        lo, hi = cuts[s], cuts[s + 1]
        p = str(tmp_path / f"sk{s}.graph")
        capi.build_slim_graph(base[lo:hi], p, M=16, ef_construction=100, labels=np.arange(lo, hi, dtype=np.uint64))
        paths.append(p)
    ix = sharding.ShardedIndex(paths, dim, device=0)
    ix.shards[0].set_ef(256)                              # the big shard is also the slow one
    for s_ in ix.shards[1:]:
        s_.set_ef(10)
    parts_l, parts_d = [], []
    for s_ in ix.shards:
        l, d = s_.search(q, k)
        parts_l.append(l)
        parts_d.append(d)
    want_l, want_d = sharding.merge_numpy(np.stack(parts_l), np.stack(parts_d), k)
    dq = torch.from_numpy(q).cuda()
    outs = [ix.search(dq, nq, k) for _ in range(6)]
    ix.join()
    for pl, pd in outs:
        assert np.array_equal(pl.cpu().numpy().view(np.uint32), want_l)
        assert np.array_equal(pd.cpu().numpy().view(np.uint32), want_d.view(np.uint32))
    ix.close()


@pytest.mark.gpu
def test_shard_group_two_ranks_one_process(tmp_path):
    """hs_shardgroup_connect_local: ONE process drives both ranks (here both on cuda:0; on a multi-GPU
    box one per device with peer access — the C++ host CLI's mode).  Each rank owns 2 of 4 shards, its
    kernels store rows into both ranks' tables; both ranks end with the merged union, batch after
    batch, with more batches in flight than the group's depth; host (pinned) buffers work in place."""
    import torch
    from hnsw_slim_b200 import capi, sharding
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k, ef, S = 16000, 500, 96, 10, 40, 4
    base, q = make_dataset(n, nq, dim, rank=8, seed=13)
    paths = _build_shards(str(tmp_path), base, S)
    shards = [capi.Index(p, dim) for p in paths]
    for s_ in shards:
        s_.set_ef(ef)
    groups = [capi.ShardGroup(shards[2 * r: 2 * r + 2], 2, r, nq, k, depth=3) for r in range(2)]
    capi.ShardGroup.connect_local(groups)
    batches = [np.ascontiguousarray(np.roll(q, b, axis=0)) for b in range(8)]
    want = []
    for qb in batches:
        pl, pd = zip(*[s_.search(qb, k) for s_ in shards])
        want.append(sharding.merge_numpy(np.stack(pl), np.stack(pd), k))
    dqs = [torch.from_numpy(b).cuda() for b in batches]
    outs = [[(torch.empty((nq, k), dtype=torch.int32, device="cuda"), torch.empty((nq, k), device="cuda"))
             for _ in batches] for _ in range(2)]
    for b in range(len(batches)):
        for r in range(2):
            groups[r].submit(dqs[b].data_ptr(), nq, outs[r][b][0].data_ptr(), outs[r][b][1].data_ptr())
    for g in groups:
        g.wait()
    for r in range(2):
        for b in range(len(batches)):
            assert np.array_equal(outs[r][b][0].cpu().numpy().view(np.uint32), want[b][0]), (r, b)
            assert np.array_equal(outs[r][b][1].cpu().numpy().view(np.uint32), want[b][1].view(np.uint32)), (r, b)
    # pinned + mapped host buffers in place (the end-to-end path of bench.py)
    hq = torch.from_numpy(batches[1]).pin_memory()
    hl = [torch.empty((nq, k), dtype=torch.int32).pin_memory() for _ in range(2)]
    hd = [torch.empty((nq, k), dtype=torch.float32).pin_memory() for _ in range(2)]
    for r in range(2):
        groups[r].submit(hq.data_ptr(), nq, hl[r].data_ptr(), hd[r].data_ptr())
    for g in groups:
        g.wait()
    for r in range(2):
        assert np.array_equal(hl[r].numpy().view(np.uint32), want[1][0])
        assert np.array_equal(hd[r].numpy().view(np.uint32), want[1][1].view(np.uint32))
    # argument errors: pageable host memory, nq above the table shape
    with pytest.raises(capi.HsError):
        groups[0].submit(batches[0].ctypes.data, nq, hl[0].data_ptr(), hd[0].data_ptr())
    with pytest.raises(capi.HsError):
        groups[0].submit(dqs[0].data_ptr(), nq + 1, outs[0][0][0].data_ptr(), outs[0][0][1].data_ptr())
    for g in groups:
        g.close()


def _build_shards(tmp, base, S, kind="slim"):
    from hnsw_slim_b200 import capi
    paths = []
    for s, (lo, hi) in enumerate(sharding.shard_ranges(len(base), S)):
        p = os.path.join(tmp, f"x{s}.graph")
        capi.build_slim_graph(base[lo:hi], p, M=16, ef_construction=100, labels=np.arange(lo, hi, dtype=np.uint64))
        paths.append(p)
    return paths


@pytest.mark.gpu
def test_scatter_search_writes_every_destination(tmp_path):
    """hs_search_batch_device_scatter: the rows of a search land in slot `slot` of every destination
    table and equal the rows of the plain device search."""
    from hnsw_slim_b200 import capi
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k = 8000, 257, 64, 10
    base, q = make_dataset(n, nq, dim, rank=8, seed=4)
    g = str(tmp_path / "g.graph")
    capi.build_slim_graph(base, g, M=16, ef_construction=100)
    ix = capi.Index(g, dim)
    ix.set_ef(50)
    want_l, want_d = ix.search(q, k)
    dq = torch.from_numpy(q).cuda()
    slots = 3
    tabs = [(torch.full((slots, nq, k), -7, dtype=torch.int32, device="cuda"),
             torch.full((slots, nq, k), -7.0, dtype=torch.float32, device="cuda")) for _ in range(2)]
    s = torch.cuda.current_stream().cuda_stream
    capi.search_device_scatter(ix, dq.data_ptr(), nq, k, [t[0].data_ptr() for t in tabs],
                               [t[1].data_ptr() for t in tabs], 1, s)
    torch.cuda.synchronize()
    for tl, td in tabs:
        assert np.array_equal(tl[1].cpu().numpy().view(np.uint32), want_l)
        assert np.array_equal(td[1].cpu().numpy().view(np.uint32), want_d.view(np.uint32))
        assert (tl[0] == -7).all() and (tl[2] == -7).all() and (td[0] == -7).all()      # other slots untouched


@pytest.mark.gpu
def test_shard_group_lagging_rank_does_not_deadlock(tmp_path):
    """One rank runs far ahead of the other (its host thread submits 12 batches at once, the other rank
    starts half a second later).  The table slot of batch s is reused by batch s + depth, whose warps wait
    in the kernel for every rank's acknowledgement of batch s: the fast rank must not fill its SMs with
    such waiting warps while its own merge of batch s still needs an SM slot (hs_shardgroup_submit holds
    the HOST until the rank's own merge of batch s is complete).  Both ranks share cuda:0 here, which
    makes the slot shortage worse than on separate GPUs."""
    import threading
    import time as _time
    import torch
    from hnsw_slim_b200 import capi, sharding
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k, ef, S = 12000, 8000, 96, 10, 30, 2
    base, q = make_dataset(n, nq, dim, rank=8, seed=19)
    paths = _build_shards(str(tmp_path), base, S)
    shards = [capi.Index(p, dim) for p in paths]
    for s_ in shards:
        s_.set_ef(ef)
    pl, pd = zip(*[s_.search(q, k) for s_ in shards])
    want_l, want_d = sharding.merge_numpy(np.stack(pl), np.stack(pd), k)
    groups = [capi.ShardGroup([shards[r]], 2, r, nq, k, depth=2) for r in range(2)]
    capi.ShardGroup.connect_local(groups)
    dq = torch.from_numpy(q).cuda()
    nb = 12
    outs = [[(torch.empty((nq, k), dtype=torch.int32, device="cuda"), torch.empty((nq, k), device="cuda"))
             for _ in range(nb)] for _ in range(2)]
    torch.cuda.synchronize()
    errs = []

    def run(r, delay):
        try:
            _time.sleep(delay)
            for b in range(nb):
                groups[r].submit(dq.data_ptr(), nq, outs[r][b][0].data_ptr(), outs[r][b][1].data_ptr())
            groups[r].wait()
        except Exception as e:      # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=run, args=(0, 0.0)), threading.Thread(target=run, args=(1, 0.5))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "shard group deadlocked"
    assert not errs, errs
    for r in range(2):
        for b in range(nb):
            assert np.array_equal(outs[r][b][0].cpu().numpy().view(np.uint32), want_l), (r, b)
            assert np.array_equal(outs[r][b][1].cpu().numpy().view(np.uint32), want_d.view(np.uint32)), (r, b)
    for g in groups:
        g.close()


def _fused_worker(rank, world, paths, qfile, dim, k, ef, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from hnsw_slim_b200 import sharding as sh
    q = np.load(qfile)
    nq = q.shape[0]
    mine = sh.shards_of_rank(len(paths), rank, world)
    ix = sh.ShardedIndex([paths[s] for s in mine], dim, device=0)
    ix.set_ef(ef)
    ix.connect(rank, world, nq, k, depth=2)
    dist.barrier()
    dqs = [torch.from_numpy(np.ascontiguousarray(np.roll(q, b, axis=0))).cuda() for b in range(3)]
    torch.cuda.synchronize()
    outs = [ix.search(dq, nq, k) for dq in dqs]         # three batches in a row: table slots are reused
    ix.join()
    res = [(l.cpu().numpy().view(np.uint32), d.cpu().numpy()) for l, d in outs]
    np.savez(os.path.join(outdir, f"r{rank}.npz"), **{f"l{b}": r[0] for b, r in enumerate(res)},
             **{f"d{b}": r[1] for b, r in enumerate(res)})
    dist.barrier()
    ix.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_fused_exchange_two_ranks_one_gpu(tmp_path):
    """hs_shardgroup over two PROCESSES (ranks) sharing cuda:0: each rank owns 2 of 4 shards, its kernels
    store result rows into BOTH ranks' gather tables (the other rank's through CUDA IPC), the last warp
    of a batch raises the flags, every rank merges on its merge stream.  Both ranks must end up with
    the (dist, label)-sorted union of the per-shard rows."""
    from hnsw_slim_b200 import capi
    from hnsw_slim_b200.synth import make_dataset
    n, nq, dim, k, ef, S = 16000, 300, 64, 10, 60, 4
    base, q = make_dataset(n, nq, dim, rank=8, seed=31)
    paths = _build_shards(str(tmp_path), base, S)
    qfile = str(tmp_path / "q.npy")
    np.save(qfile, q)
    want = []
    for b in range(3):
        qb = np.ascontiguousarray(np.roll(q, b, axis=0))
        pl, pd = [], []
        for p in paths:
            one = capi.Index(p, dim)
            one.set_ef(ef)
            l, d = one.search(qb, k)
            pl.append(l)
            pd.append(d)
        want.append(sharding.merge_numpy(np.stack(pl), np.stack(pd), k))
    port = 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, paths, qfile, dim, k, ef, port, str(tmp_path)))
             for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "fused-exchange ranks hung"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        for b in range(3):
            assert np.array_equal(z[f"l{b}"], want[b][0]), (r, b)
            assert np.array_equal(z[f"d{b}"].view(np.uint32), want[b][1].view(np.uint32)), (r, b)
