"""One-off: the bench graph is built by the engine's host builder; does the CPU arm — and the
GPU — behave the same on a graph built by the REFERENCE (HierarchicalNSW::addPoint + convertFromHNSW, OpenMP)?
Builds both graphs over the C1 corpus and reports recall / QPS of both arms on both."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnsw_slim_b200 import capi  # noqa: E402
from hnsw_slim_b200.synth import latent_gaussian  # noqa: E402
from oracle import refharness as rh  # noqa: E402

n, dim, M, efc, ef, k, nq = 1_000_000, 128, 16, 200, 100, 10, 10_000
base = latent_gaussian(n, dim, rank=14, seed=1)
q = latent_gaussian(nq, dim, rank=14, seed=1, stream=1)
gt, _ = capi.bruteforce_knn(base, q[:2000], k)
out = {"cores": os.cpu_count()}
paths = {}
t0 = time.time()
capi.build_slim_graph(base, "/tmp/c1_engine.graph", M=M, ef_construction=efc, branching="4")
out["engine_build_s"] = time.time() - t0
paths["engine_host_builder"] = "/tmp/c1_engine.graph"
t0 = time.time()
rh.ref_slim_build(base, "/tmp/c1_ref.graph", M=M, ef_construction=efc, branching="4")
out["reference_build_s"] = time.time() - t0
paths["reference_builder"] = "/tmp/c1_ref.graph"
t0 = time.time()
gix = capi.Index.build_gpu(base, M=M, ef_construction=efc)
out["gpu_build_s"] = time.time() - t0
gix.save("/tmp/c1_gpu.graph")
paths["engine_gpu_builder"] = "/tmp/c1_gpu.graph"
for name, p in paths.items():
    ix = capi.Index(p, dim)
    ix.set_ef(ef)
    ix.set_overlap(True)
    info = ix.info()
    lab, _ = ix.search(q[:2000], k)
    rec_gpu = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(lab, gt)]))
    import torch
    dq = torch.from_numpy(q).cuda()
    dl = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    dd = torch.empty((nq, k), device="cuda")
    st = torch.cuda.Stream()
    for _ in range(5):
        ix.search_device(dq.data_ptr(), nq, k, dl.data_ptr(), dd.data_ptr(), st.cuda_stream)
    st.synchronize()
    ix.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50):
        ix.search_device(dq.data_ptr(), nq, k, dl.data_ptr(), dd.data_ptr(), st.cuda_stream)
    e1.record(st)
    st.synchronize()
    gpu_qps = nq * 50 / (e0.elapsed_time(e1) * 1e-3)
    stats = ix.stats()
    ref = rh.RefSlim(p, dim, n, 0)
    ref.search(q[:2000], k, ef, 0)
    times = []
    rlab = None
    while sum(times) < 8.0:
        rlab, sec, _ = ref.search(q, k, ef, 0)
        times.append(sec)
    rec_cpu = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(rlab[:2000], gt)]))
    out[name] = {"avg_deg0": info["sum_deg0"] / n, "max_deg0": info["max_deg0"], "maxlevel": info["maxlevel"],
                 "gpu_recall": rec_gpu, "gpu_qps": gpu_qps, "gpu_evals_per_query": stats["n_dist"] / (50 * nq),
                 "cpu_recall": rec_cpu, "cpu_qps_all_cores": nq / float(np.median(times)), "ratio": gpu_qps / (nq / float(np.median(times)))}
    print(name, json.dumps(out[name]), flush=True)
    ref.close()
print(json.dumps(out))
