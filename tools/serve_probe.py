"""Developer probe for the serving + delta-patch path, run on a GPU box:
the reference's server side (oracle/_ref: addPoint + convertFromHNSWWithDiff) writes a partial index and the patch
streams, hs_main --solve_strategy=hnsw_slim_serve applies them to the HBM-resident index and answers the query set
as single queries from T concurrent callers through hs_service.  Prints patch-apply times and QPS / batch shape /
recall per T, next to the one-call batched search on the server's final index."""
import argparse
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnsw_slim_b200 import build as hs_build  # noqa: E402
from hnsw_slim_b200 import vecs_io  # noqa: E402
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--n0", type=int, default=190000)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--nq", type=int, default=100000)
ap.add_argument("--ef", type=int, default=100)
ap.add_argument("--threads", type=str, default="16,64,256,1024,4096")
a = ap.parse_args()

td = tempfile.mkdtemp(prefix="hs_serve_")
base, q = make_dataset(a.n, a.nq, a.dim, rank=14, seed=1)
d = os.path.join(td, "data", "toy")
os.makedirs(d)
vecs_io.write_vecs(os.path.join(d, "toy_base.fvecs"), base)
vecs_io.write_vecs(os.path.join(d, "toy_query.fvecs"), q)
part, fin = os.path.join(td, "part.graph"), os.path.join(td, "final.graph")
t0 = time.time()
names = rh.ref_slim_make_patches(base, a.n0, a.rounds, part, os.path.join(td, "p"), final_path=fin, inline_last=True,
                                 M=16, ef_construction=200, threads=0)
print(f"reference server side: partial index of {a.n0} rows + {a.rounds} updates to {a.n} rows in {time.time()-t0:.1f} s; "
      f"patch streams {[os.path.getsize(x) for x in names]} bytes", flush=True)
common = ["--dataset=toy", "--data_dir", os.path.join(td, "data"), "--index_dir", os.path.join(td, "index"), "--k=10",
          f"--ef_search={a.ef}"]


def cli(*args):
    r = subprocess.run([hs_build.HOST_CLI, *args], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
    return r.stdout


cli("--solve_strategy=bruteforce", *common)
out = cli("--solve_strategy=hnsw_slim", *common, "--index_path", fin)
ms = float(re.search(r"solve cost: ([0-9.]+)", out).group(1))
print(f"one batched call on the server's final index: {ms:.2f} ms for {a.nq} queries = {a.nq/ms*1e3:,.0f} QPS, "
      f"recall {re.search(r'Recall: ([0-9.]+)', out).group(1)}", flush=True)
for t in [int(x) for x in a.threads.split(",")]:
    out = cli("--solve_strategy=hnsw_slim_serve", *common, "--index_path", part, "--patches", ",".join(names),
              "--patch_inline_last", f"--threads={t}", "--max_batch=4096")
    if t == int(a.threads.split(",")[0]):
        for line in out.splitlines():
            if line.startswith("patch "):
                print("  " + line, flush=True)
    ms = float(re.search(r"solve cost: ([0-9.]+)", out).group(1))
    served = re.search(r"served .*", out).group(0)
    print(f"threads={t:4d}: {ms:9.2f} ms = {a.nq/ms*1e3:11,.0f} QPS; {served}; recall {re.search(r'Recall: ([0-9.]+)', out).group(1)}",
          flush=True)
