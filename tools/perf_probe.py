"""Developer probe (not the bench contract): time hs_search_batch_device on a synthetic corpus.
Graph comes from the reference builder via oracle/_ref (dev tool; bench.py does not do this)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnsw_slim_b200 import capi  # noqa: E402
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--rank", type=int, default=14)
ap.add_argument("--M", type=int, default=16)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--metric", type=int, default=0)
ap.add_argument("--efs", type=str, default="50,100,200")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()

base, q = make_dataset(a.n, a.nq, a.dim, metric=a.metric, rank=a.rank)
cache = os.environ.get("HS_DATA_CACHE", "/tmp/hs_data_cache")
os.makedirs(cache, exist_ok=True)
graph = os.path.join(cache, f"probe_n{a.n}_d{a.dim}_r{a.rank}_M{a.M}_e{a.efc}_m{a.metric}.graph")
if not os.path.exists(graph):
    t = time.time()
    bs, cs = rh.ref_slim_build(base, graph, metric=a.metric, M=a.M, ef_construction=a.efc, branching="4")
    print(f"built graph in {time.time()-t:.1f}s (addPoint {bs:.1f}s, convert {cs:.1f}s) with {os.cpu_count()} cpus", flush=True)
t = time.time()
ix = capi.Index(graph, a.dim, metric=a.metric)
info = ix.info()
print(f"load {time.time()-t:.1f}s", {k: info[k] for k in ("n", "dim_padded", "maxlevel", "deg0_stride", "max_deg0", "upper_stride", "n_upper", "sum_deg0", "device_bytes")}, flush=True)

dq = torch.from_numpy(q).cuda()
dl = torch.empty((a.nq, a.k), dtype=torch.int32, device="cuda")
dd = torch.empty((a.nq, a.k), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream()
gt = None
if a.check:
    gl, gd = capi.bruteforce_knn(base, q[:1000], a.k, metric=a.metric)
    gt = [set(r) for r in gl]
for ef in [int(x) for x in a.efs.split(",")]:
    ix.set_ef(ef)
    for _ in range(2):
        ix.search_device(dq.data_ptr(), a.nq, a.k, dl.data_ptr(), dd.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    ix.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters):
        ix.search_device(dq.data_ptr(), a.nq, a.k, dl.data_ptr(), dd.data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    st = ix.stats()
    nd, nh = st["n_dist"] / a.iters / a.nq, st["n_hops"] / a.iters / a.nq
    bytes_q = nd * 4 * info["dim_padded"] + nh * (8 + 4 * info["sum_deg0"] / info["n"])
    rec = -1.0
    if gt is not None:
        lab = dl[:1000].cpu().numpy().view(np.uint32)
        rec = float(np.mean([len(set(r) & g) / a.k for r, g in zip(lab, gt)]))
    print(f"ef={ef:4d}  {ms:8.3f} ms/batch  {a.nq/ms*1e3:12.0f} QPS  n_dist/q={nd:8.1f} n_hops/q={nh:6.1f} "
          f"alg GB/s={bytes_q*a.nq/ms/1e6:8.1f}  recall@{a.k}={rec:.4f}", flush=True)
