"""Times hs_build_slim_index_gpu and compares the graph with the host builder's (recall / evaluations)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnsw_slim_b200 import capi  # noqa: E402
from hnsw_slim_b200.synth import latent_gaussian_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=96)
ap.add_argument("--rank", type=int, default=12)
ap.add_argument("--M", type=int, default=16)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--metric", type=int, default=0)
ap.add_argument("--host", action="store_true", help="also build with the host builder and compare")
ap.add_argument("--efs", default="16,24,50,100")
a = ap.parse_args()
t0 = time.time()
base = latent_gaussian_rows(0, a.n, a.dim, rank=a.rank, seed=1, normalize=a.metric == 1)
q = latent_gaussian_rows(0, 2000, a.dim, rank=a.rank, seed=1, stream=1, normalize=a.metric == 1)
print(f"data {time.time()-t0:.1f}s", flush=True)
t0 = time.time()
gpu = capi.Index.build_gpu(base, metric=a.metric, M=a.M, ef_construction=a.efc)
tb = time.time() - t0
gi = gpu.info()
print(f"GPU build n={a.n} dim={a.dim} M={a.M} efc={a.efc}: {tb:.2f}s ({a.n/tb/1e3:.0f}k rows/s), avg deg0 {gi['sum_deg0']/a.n:.2f}, "
      f"max deg0 {gi['max_deg0']}, maxlevel {gi['maxlevel']}", flush=True)
gt, _ = capi.bruteforce_knn(base, q, 10, metric=a.metric)
idx = {"gpu": gpu}
if a.host:
    t0 = time.time()
    capi.build_slim_graph(base, "/tmp/probe_host.graph", metric=a.metric, M=a.M, ef_construction=a.efc)
    print(f"host build: {time.time()-t0:.1f}s ({os.cpu_count()} threads)", flush=True)
    idx["host"] = capi.Index("/tmp/probe_host.graph", a.dim, metric=a.metric)
    hi = idx["host"].info()
    print(f"host avg deg0 {hi['sum_deg0']/a.n:.2f} max {hi['max_deg0']}")
for ef in [int(x) for x in a.efs.split(",")]:
    for name, ix in idx.items():
        ix.set_ef(ef)
        ix.reset_stats()
        lab, _ = ix.search(q, 10)
        rec = np.mean([len(set(x) & set(y)) / 10 for x, y in zip(lab, gt)])
        print(f"ef={ef:4d} {name:5s} recall {rec:.4f} evals/query {ix.stats()['n_dist']/len(q):.0f}", flush=True)
