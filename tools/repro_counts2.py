"""Repro for a rare counter mismatch: the test's own sequence (new index per ef, ef cycling),
repeated; on a mismatch print the query, both counters and the counters the previous ef gave."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import get_corpus  # noqa: E402
from hnsw_slim_b200 import capi  # noqa: E402
from oracle import refharness as rh  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
efs = (10, 60, 150, 300)
want = {}
for thr in (1, 2, 7):
    c = get_corpus(n=20000, nq=300, dim=32, threshold_level=thr)
    orc = rh.Oracle(c.graph, c.dim, c.metric)
    for ef in efs:
        want[thr, ef] = orc.search(c.queries, 10, ef, order=rh.ORDER_GPU, team=8)
bad = 0
for r in range(reps):
    for thr in (1, 2, 7):
        c = get_corpus(n=20000, nq=300, dim=32, threshold_level=thr)
        prev = None
        for ef in efs:
            ix = capi.Index(c.graph, c.dim, metric=c.metric)
            ix.set_ef(ef)
            lab, dist, cnt = ix.search(c.queries, 10, counts=True)
            ol, od, ond, onh = want[thr, ef]
            same = np.all(lab == ol, axis=1)
            mism = np.nonzero(same & ((cnt[:, 0] != ond) | (cnt[:, 1] != onh)))[0]
            rowm = np.nonzero(~same)[0]
            if len(mism) or len(rowm):
                bad += 1
                for q in list(mism[:4]) + list(rowm[:2]):
                    print(f"rep={r} thr={thr} ef={ef} query={q} gpu counts={cnt[q]} oracle=({ond[q]},{onh[q]}) "
                          f"prev-ef gpu counts={None if prev is None else prev[q]} rows_equal={bool(same[q])}", flush=True)
            prev = cnt
print(f"{bad} mismatching (thr, ef) runs out of {reps * 12}")
