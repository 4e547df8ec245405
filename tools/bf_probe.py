"""Developer probe: exact-kNN (hs_bruteforce_knn_device) timing, tcgen05 path vs fp32 scan kernel."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnsw_slim_b200 import capi  # noqa: E402
from hnsw_slim_b200.synth import make_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--paths", default="1,0")
a = ap.parse_args()
base, q = make_dataset(a.n, a.nq, a.dim, metric=a.metric, rank=14)
db, dq = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
dl = torch.empty((a.nq, a.k), dtype=torch.int32, device="cuda")
dd = torch.empty((a.nq, a.k), dtype=torch.float32, device="cuda")
res = {}
for tc in a.paths.split(","):
    os.environ["HS_BF_TC"] = tc
    os.environ["HS_BF_TC_STATS"] = "1"
    s = torch.cuda.current_stream().cuda_stream
    capi.bruteforce_knn_device(db.data_ptr(), a.n, a.dim, dq.data_ptr(), a.nq, a.k, dl.data_ptr(), dd.data_ptr(),
                               metric=a.metric, stream=s)
    torch.cuda.synchronize()
    fb = capi.bf_tc_fallback()
    del os.environ["HS_BF_TC_STATS"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        capi.bruteforce_knn_device(db.data_ptr(), a.n, a.dim, dq.data_ptr(), a.nq, a.k, dl.data_ptr(), dd.data_ptr(),
                                   metric=a.metric, stream=s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    flop = 2.0 * a.n * a.nq * a.dim
    res[tc] = (dl.cpu().numpy().copy(), dd.cpu().numpy().copy())
    print(f"path={'tcgen05' if tc == '1' else 'scan   '} n={a.n} nq={a.nq} dim={a.dim} k={a.k} metric={a.metric}: "
          f"{ms:9.2f} ms  {flop/ms/1e9:8.1f} TFLOP/s (algorithmic 2*n*nq*dim)  fallback={fb}", flush=True)
if len(res) == 2:
    (l1, d1), (l0, d0) = res["1"], res["0"]
    print("identical ids:", np.array_equal(l0, l1), " identical distances:", np.array_equal(d0.view(np.uint32), d1.view(np.uint32)))
