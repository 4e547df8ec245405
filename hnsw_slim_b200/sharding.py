"""Single-box partitioning for corpora that are sharded into per-GPU sub-graphs.

The reference builds ONE graph even for its 100M datasets; here the base set is split into S
contiguous label ranges, one HNSW-Slim sub-graph per range (labels stay global), the S shards are
spread over the ranks (S/N per GPU), every rank searches the whole query batch on its shards, and
one all-gather of nq x k x 8 bytes per rank plus the top-k merge kernel produce the global result
on every rank.  1M-scale indices do not need any of this: they are replicated and the queries are
split (bench.py --gpus N).

The collective goes through torch.distributed (NCCL on GPUs; gloo in the CPU tests, which inject a
numpy merge because the product merge is a CUDA kernel).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def shard_ranges(n: int, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous [begin, end) label ranges, sizes differing by at most one."""
    base, rem = divmod(n, n_shards)
    out, b = [], 0
    for s in range(n_shards):
        e = b + base + (1 if s < rem else 0)
        out.append((b, e))
        b = e
    return out


def shards_of_rank(n_shards: int, rank: int, world: int) -> list[int]:
    """Round-robin-free block assignment: rank r owns shards [r*S/N, (r+1)*S/N)."""
    if n_shards % world != 0:
        raise ValueError(f"{n_shards} shards do not divide over {world} ranks")
    per = n_shards // world
    return list(range(rank * per, (rank + 1) * per))


def merge_numpy(labels: np.ndarray, dists: np.ndarray, k: int):
    """Reference semantics of hs_topk_merge_device: [parts, nq, k] -> [nq, k] by (dist, label);
    label 0xFFFFFFFF marks padding."""
    parts, nq, kk = labels.shape
    lab = labels.transpose(1, 0, 2).reshape(nq, parts * kk)
    dst = dists.transpose(1, 0, 2).reshape(nq, parts * kk).astype(np.float32)
    out_l = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    for i in range(nq):
        ok = lab[i] != 0xFFFFFFFF
        order = np.lexsort((lab[i][ok], dst[i][ok]))[:k]
        out_l[i, : len(order)] = lab[i][ok][order]
        out_d[i, : len(order)] = dst[i][ok][order]
    return out_l, out_d


def gather_and_merge(local_labels, local_dists, k: int, *, group=None, merge: Callable | None = None):
    """All-gather every rank's [nq, k] partial result and merge to the global top-k.

    local_labels (int32/uint32 view) and local_dists (float32) are torch tensors on the device of
    the process group's backend.  `merge(labels[parts,nq,k], dists[parts,nq,k], k)` defaults to the
    CUDA merge kernel (hs_topk_merge_device)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = local_labels.shape[0]
    if world > 1:
        # concatenated along dim 0 (the layout gloo and NCCL both accept), viewed as [world, nq, k]
        all_l = torch.empty((world * nq, k), dtype=local_labels.dtype, device=local_labels.device)
        all_d = torch.empty((world * nq, k), dtype=local_dists.dtype, device=local_dists.device)
        dist.all_gather_into_tensor(all_l, local_labels.contiguous(), group=group)
        dist.all_gather_into_tensor(all_d, local_dists.contiguous(), group=group)
        all_l, all_d = all_l.view(world, nq, k), all_d.view(world, nq, k)
    else:
        all_l, all_d = local_labels.unsqueeze(0), local_dists.unsqueeze(0)
    if merge is not None:
        return merge(all_l, all_d, k)
    from . import capi
    out_l = torch.empty((nq, k), dtype=local_labels.dtype, device=local_labels.device)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=local_labels.device)
    capi.topk_merge_device(all_l.data_ptr(), all_d.data_ptr(), world, nq, k, out_l.data_ptr(), out_d.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    return out_l, out_d


class ShardedIndex:
    """The shards one rank owns, resident on its GPU; search() returns the GLOBAL top-k."""

    def __init__(self, graph_paths: Sequence[str], dim: int, *, metric: int = 0, device: int = 0,
                 kind: int = 0, raw_bases: Sequence | None = None):
        """kind = capi.HS_KIND_SLIMQ needs raw_bases[i] = the rows of shard i (exact rerank, indexed
        by the shard's internal ids)."""
        from . import capi
        self.capi = capi
        if kind == capi.HS_KIND_SLIMQ:
            assert raw_bases is not None and len(raw_bases) == len(graph_paths)
            self.shards = [capi.Index(p, dim, kind=kind, raw_base=rb, device=device)
                           for p, rb in zip(graph_paths, raw_bases)]
        else:
            self.shards = [capi.Index(p, dim, metric=metric, device=device) for p in graph_paths]
        # the shard launches of one batch are independent (same queries, own outputs): let each one's
        # tail overlap the next one's head on the stream (hs_set_overlap)
        for s in self.shards:
            s.set_overlap(True)
        self.dim = dim
        self._side = None                    # stream of the pipelined exchange step (search(pipelined=True))
        self._ex = None                      # fused exchange (connect_exchange / search_fused)

    def set_ef(self, ef: int) -> None:
        for s in self.shards:
            s.set_ef(ef)

    def device_bytes(self) -> int:
        return sum(s.info()["device_bytes"] for s in self.shards)

    def search_local(self, d_queries, nq: int, k: int):
        """Every local shard on the whole batch, then the local top-k (device tensors)."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        n_loc = len(self.shards)
        lab = torch.empty((n_loc, nq, k), dtype=torch.int32, device=d_queries.device)
        dst = torch.empty((n_loc, nq, k), dtype=torch.float32, device=d_queries.device)
        for i, s in enumerate(self.shards):
            s.search_device(d_queries.data_ptr(), nq, k, lab[i].data_ptr(), dst[i].data_ptr(), stream)
        if n_loc == 1:
            return lab[0], dst[0]
        out_l = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
        self.capi.topk_merge_device(lab.data_ptr(), dst.data_ptr(), n_loc, nq, k, out_l.data_ptr(),
                                    out_d.data_ptr(), stream)
        return out_l, out_d

    def search(self, d_queries, nq: int, k: int, group=None, pipelined: bool = False):
        """Global top-k of one batch.  pipelined=True runs the exchange step (all-gather + merge) on
        a side stream, so the NEXT batch's shard searches — enqueued on the caller's stream — do
        not wait for it: a stream of batches then costs max(search, exchange) per batch instead of
        their sum — in principle: measured on 2 B200s it is slower (2.43 M vs 2.68 M QPS), because the
        NCCL kernels queue behind the persistent traversal grid that owns every SM slot, so it is off
        by default.  The returned tensors are complete once join() (or a device synchronize) has
        been called."""
        l, d = self.search_local(d_queries, nq, k)
        if not pipelined:
            return gather_and_merge(l, d, k, group=group)
        import torch
        if self._side is None:
            self._side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            out = gather_and_merge(l, d, k, group=group)
        l.record_stream(self._side)          # allocated on the caller's stream, consumed on the side stream
        d.record_stream(self._side)
        return out

    # ---- exchange fused into the traversal kernel (no collective) ----
    def connect_exchange(self, rank: int, world: int, n_shards_total: int, nq_max: int, k: int, device: int,
                         group=None):
        """Create this rank's gather tables and map every other rank's (CUDA IPC; the 64-byte handles
        travel through torch.distributed).  Call once, collectively."""
        import torch.distributed as dist
        ex = self.capi.Exchange(device, world, rank, n_shards_total, nq_max, k)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, ex.handle(), group=group)
            ex.connect(b"".join(handles))
        self._ex, self._ex_rank, self._ex_world, self._ex_slots, self._ex_seq = ex, rank, world, n_shards_total, 0
        return ex

    def search_fused(self, d_queries, nq: int, k: int):
        """Global top-k of one batch with the exchange fused into the traversal kernels: every local
        shard's kernel stores its rows into slot (rank * shards_per_rank + i) of EVERY rank's gather
        table (peer memory over NVLink), a stream-ordered flag exchange tells when all rows of the
        batch have landed, and hs_topk_merge_device reads the local table.  No all-gather call."""
        import torch
        ex = self._ex
        self._ex_seq += 1
        seq = self._ex_seq
        stream = torch.cuda.current_stream().cuda_stream
        n_loc = len(self.shards)
        for i, s in enumerate(self.shards):
            ex.search(s, d_queries.data_ptr(), nq, self._ex_rank * n_loc + i, seq, stream)
        ex.signal_and_wait(seq, stream)
        tl, td = ex.tables(seq)
        out_l = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
        self.capi.topk_merge_device(tl, td, self._ex_slots, nq, k, out_l.data_ptr(), out_d.data_ptr(), stream)
        return out_l, out_d

    def join(self) -> None:
        """Make the caller's stream wait for every pipelined exchange issued so far."""
        import torch
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
