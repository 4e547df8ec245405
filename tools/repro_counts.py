"""Determinism probe: repeat a search many times and report queries whose counters or results change."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import get_corpus  # noqa: E402
from hnsw_slim_b200 import capi  # noqa: E402

thr = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
c = get_corpus(n=20000, nq=300, dim=32, threshold_level=thr) if thr else get_corpus(n=20000, nq=300, dim=32)
for ef in (60, 150, 300):
    ix = capi.Index(c.graph, c.dim)
    ix.set_ef(ef)
    l0, d0, c0 = ix.search(c.queries, 10, counts=True)
    bad = 0
    for r in range(reps):
        ix = capi.Index(c.graph, c.dim)
        ix.set_ef(ef)
        l, d, cn = ix.search(c.queries, 10, counts=True)
        diff = np.nonzero((cn != c0).any(1) | (l != l0).any(1))[0]
        if len(diff):
            bad += 1
            for q in diff[:3]:
                print(f"thr={thr} ef={ef} rep={r} query={q} counts {c0[q]} -> {cn[q]} labels_equal={np.array_equal(l[q], l0[q])}")
    print(f"thr={thr} ef={ef} ghash={os.environ.get('HS_GHASH','auto')}: {bad}/{reps} runs differ from the first", flush=True)
