"""Single-box partitioning for corpora that are sharded into per-GPU sub-graphs.

The reference builds ONE graph even for its 100M datasets; here the base set is split into S
contiguous label ranges, one HNSW-Slim sub-graph per range (labels stay global), the S shards are
spread over the ranks (S/N per GPU), every rank searches the whole query batch on its shards, the
traversal kernels store their rows into the gather tables of every rank (hs_shardgroup; or one
all-gather of nq x k x 8 bytes per rank) and a top-k merge kernel produces the global result on
every rank.  1M-scale indices do not need any of this: they are replicated and the queries are
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
    """All-gather every rank's [nq, k] partial result and merge to the global top-k (the NCCL form of
    the exchange; the default on GPUs is the exchange fused into the traversal kernels, ShardedIndex).

    local_labels (int32/uint32 view) and local_dists (float32) are torch tensors on the device of
    the process group's backend; both travel in ONE all-gather (labels and the distances' bit
    patterns stacked into a [2, nq, k] int32 block).  `merge(labels[parts,nq,k], dists[parts,nq,k], k)`
    defaults to the CUDA merge kernel (hs_topk_merge_device)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = local_labels.shape[0]
    if world > 1:
        both = torch.stack((local_labels.contiguous().view(torch.int32),
                            local_dists.contiguous().view(torch.int32)))              # [2, nq, k]
        gathered = torch.empty((world * 2, nq, k), dtype=torch.int32, device=both.device)
        dist.all_gather_into_tensor(gathered, both, group=group)
        gathered = gathered.view(world, 2, nq, k)
        all_l = gathered[:, 0].contiguous().view(local_labels.dtype)
        all_d = gathered[:, 1].contiguous().view(torch.float32)
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
    """The shards one rank owns, resident on its GPU.  search() returns the GLOBAL top-k.

    Everything on the GPU goes through one hs_shardgroup (csrc/shard_group.cu): the traversal kernels
    store their rows into every rank's gather table (peer memory), the last finishing warp of a
    batch raises the flags, a merge kernel on a second stream produces the result — no collective and
    nothing between two traversal launches, so a stream of batches is pipelined.  `exchange="nccl"`
    keeps the all-gather form (local group of world size 1, then gather_and_merge)."""

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
        self.dim = dim
        self.device = device
        self.group = None                    # hs_shardgroup over all ranks (connect)
        self._local = None                   # hs_shardgroup of world size 1 (NCCL form / single rank)
        self._local_shape = None

    @classmethod
    def from_indices(cls, shards, dim: int, *, device: int = 0) -> "ShardedIndex":
        """Shards that are already resident (e.g. built on the GPU, capi.Index.build_gpu)."""
        from . import capi
        self = cls.__new__(cls)
        self.capi = capi
        self.shards = list(shards)
        self.dim, self.device = dim, device
        self.group = self._local = self._local_shape = None
        return self

    def set_ef(self, ef: int) -> None:
        for s in self.shards:
            s.set_ef(ef)

    def device_bytes(self) -> int:
        return sum(s.info()["device_bytes"] for s in self.shards)

    # ---- the pipelined group over all ranks ----
    def connect(self, rank: int, world: int, nq_max: int, k: int, *, depth: int = 4, group=None):
        """Create this rank's gather tables and map every other rank's (CUDA IPC; the 64-byte handles
        travel through torch.distributed).  Collective.  Every rank reaches the all-gather even when
        its own setup failed, so that a failure raises everywhere instead of hanging the others."""
        err, g, h = None, None, None
        try:
            g = self.capi.ShardGroup(self.shards, world, rank, nq_max, k, depth)
            h = g.handle() if world > 1 else None
        except Exception as e:            # reported after the collective
            err = e
        if world > 1:
            import torch.distributed as dist
            handles = [None] * world
            dist.all_gather_object(handles, h, group=group)
            if err is None and any(x is None for x in handles):
                err = RuntimeError("another rank could not create its shard group")
            if err is None:
                try:
                    g.connect(b"".join(handles))
                except Exception as e:
                    err = e
            oks = [None] * world
            dist.all_gather_object(oks, err is None, group=group)
            if err is None and not all(oks):
                err = RuntimeError("another rank could not map the peer tables")
        elif err is None:
            g.connect(None)
        if err is not None:
            if g is not None:
                g.close()
            raise err
        self.group = g
        return g

    def submit(self, q_ptr: int, nq: int, lab_ptr: int, dist_ptr: int | None) -> None:
        """hs_shardgroup_submit on raw pointers (device, or pinned + mapped host memory)."""
        self.group.submit(q_ptr, nq, lab_ptr, dist_ptr)

    def _local_group(self, nq: int, k: int):
        if self._local is None or self._local_shape != (nq, k):
            if self._local is not None:
                self._local.wait()
                self._local.close()
            self._local = self.capi.ShardGroup(self.shards, 1, 0, nq, k, 4)
            self._local.connect(None)
            self._local_shape = (nq, k)
        return self._local

    def search_local(self, d_queries, nq: int, k: int):
        """Every local shard on the whole batch, then the local top-k (device tensors, complete after
        join()).  Single-rank form of the pipelined group."""
        import torch
        g = self._local_group(nq, k)
        out_l = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
        g.submit(d_queries.data_ptr(), nq, out_l.data_ptr(), out_d.data_ptr())
        return out_l, out_d

    def search(self, d_queries, nq: int, k: int, group=None, exchange: str = "fused"):
        """Global top-k of one batch (device tensors).  The call returns once the batch is enqueued; the
        tensors are complete after join().  exchange="fused": the connected hs_shardgroup (or the local
        one when this is the only rank); "nccl": local shards + one all-gather + hs_topk_merge_device."""
        import torch
        if exchange == "fused":
            if self.group is None:
                return self.search_local(d_queries, nq, k)
            out_l = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
            self.group.submit(d_queries.data_ptr(), nq, out_l.data_ptr(), out_d.data_ptr())
            return out_l, out_d
        l, d = self.search_local(d_queries, nq, k)
        self._local.wait()                   # the all-gather reads the local result on torch's stream
        return gather_and_merge(l, d, k, group=group)

    def join(self) -> None:
        """Block until every batch enqueued so far is complete."""
        if self.group is not None:
            self.group.wait()
        if self._local is not None:
            self._local.wait()

    def close(self) -> None:
        for g in (self.group, self._local):
            if g is not None:
                g.wait()
                g.close()
        self.group = self._local = None
