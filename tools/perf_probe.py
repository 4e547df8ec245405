"""Developer probe (not the bench contract): time hs_search_batch_device over an ef sweep on a
bench.py workload; prints QPS, counters, algorithmic GB/s and recall."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hnsw_slim_b200 import capi  # noqa: E402
from hnsw_slim_b200.synth import latent_gaussian  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sift1m")
ap.add_argument("--efs", type=str, default="100")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--check", type=int, default=0)
ap.add_argument("--nq", type=int, default=0)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--tflags", type=str, default="", help="comma list of hs_set_tuning('traverse_flags') values to sweep")
ap.add_argument("--modes", type=str, default="", help="comma list of hs_set_tuning('visited_table') values to sweep")
ap.add_argument("--gpu-build", type=int, default=0, help="build the index on the GPU instead of loading a host-built .graph")
a = ap.parse_args()
w = dict(bench.WORKLOADS[a.workload])
if a.nq:
    w["nq"] = a.nq
slimq = w.get("kind") == "slimq"
if a.gpu_build:
    base = latent_gaussian(w["n"], w["dim"], rank=w["rank"], seed=1, normalize=(w["metric"] == 1))
    qb = [latent_gaussian(w["nq"], w["dim"], rank=w["rank"], seed=1, normalize=(w["metric"] == 1), stream=1 + b) for b in range(4)]
    ix = capi.Index.build_gpu(base, metric=w["metric"], M=w["M"], ef_construction=w["efc"],
                              kind=capi.HS_KIND_SLIMQ if slimq else capi.HS_KIND_SLIM)
else:
    graph, base, qb = bench.prepare_inputs(w, 4, True)
if a.gpu_build:
    pass
elif slimq:
    ix = capi.Index(graph, w["dim"], kind=capi.HS_KIND_SLIMQ, raw_base=base)
else:
    ix = capi.Index(graph, w["dim"], metric=w["metric"])
ix.set_overlap(bool(a.overlap))
info = ix.info()
nq, k = w["nq"], w["k"]
dq = [torch.from_numpy(q).cuda() for q in qb]
dl = torch.empty((nq, k), dtype=torch.int32, device="cuda")
dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream()
gt = None
if a.check:
    if base is None:
        base = latent_gaussian(w["n"], w["dim"], rank=w["rank"], seed=1, normalize=(w["metric"] == 1))
    gl, gd = capi.bruteforce_knn(base, qb[0][:1000], k, metric=w["metric"])
    gt = [set(r) for r in gl]
modes = [int(x) for x in a.modes.split(",")] if a.modes else [None]
tflags = [int(x) for x in a.tflags.split(",")] if a.tflags else [None]
for tf, mode, ef in [(t, m, int(x)) for t in tflags for m in modes for x in a.efs.split(",")]:
    if mode is not None:
        ix.set_tuning("visited_table", mode)
    if tf is not None:
        ix.set_tuning("traverse_flags", tf)
    ix.set_ef(ef)
    for i in range(3):
        ix.search_device(dq[i % 4].data_ptr(), nq, k, dl.data_ptr(), dd.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    ix.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(a.iters):
        ix.search_device(dq[i % 4].data_ptr(), nq, k, dl.data_ptr(), dd.data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    st = ix.stats()
    nd, nh = st["n_dist"] / a.iters / nq, st["n_hops"] / a.iters / nq
    if slimq:
        nr = st["n_rerank"] / a.iters / nq
        bytes_q = (nd * (info["padded_dim_q"] // 8 + 16) + nr * 4 * info["dim_padded"]
                   + nh * (8 + 4 * info["sum_deg0"] / info["n"]))
    else:
        bytes_q = nd * 4 * info["dim_padded"] + nh * (8 + 4 * info["sum_deg0"] / info["n"])
    rec = -1.0
    if gt is not None:
        ix.search_device(dq[0].data_ptr(), nq, k, dl.data_ptr(), dd.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        lab = dl[:1000].cpu().numpy().view(np.uint32)
        rec = float(np.mean([len(set(r) & g) / k for r, g in zip(lab, gt)]))
    print(f"overlap={a.overlap} lib={os.path.basename(os.path.dirname(os.environ.get('HS_LIB_PATH','default/x')))} "
          f"flags={tf if tf is not None else os.environ.get('HS_TRAVERSE_FLAGS','-')} hb={os.environ.get('HS_HASH_BITS','-')} vt={mode} ef={ef:4d} "
          f"{ms:8.3f} ms/batch {nq/ms*1e3:11.0f} QPS  n_dist/q={nd:7.1f} n_hops/q={nh:6.1f} "
          f"alg GB/s={bytes_q*nq/ms/1e6:7.1f} recall@{k}={rec:.4f}", flush=True)
