"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref, the
unmodified hnswlib fork compiled by oracle/Makefile).  Run in the build container, where
/root/reference exists:

    python tests/golden/make_golden.py

Outputs (small, committed):
  slim_l2_2k.graph        .graph written by the reference's saveIndex for 2000 x 16 L2 vectors (M=8)
  slim_ip_1k.graph        same for 1000 x 16 unit vectors, InnerProductSpace (M=8)
  slim_l2_2k.npz / slim_ip_1k.npz:
      queries             the query batch
      ref_labels_ef{E}    the reference's searchKnn(q, k=10) labels per ef (unordered rows)
      ref_counts_ef{E}    distance evaluations per query counted inside the reference's DISTFUNC
      ref_gt100           BruteForce::solve rows (k=100, farthest first)
      ref_dist_samples    reference DISTFUNC values for (query i, base row i) pairs
      node_*              level / label / level-l neighbour slices of sampled nodes, via the
                          reference's own accessors
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EFS = (10, 40, 100)


def make(name, n, dim, metric, M, nq=64, k=10):
    base, q = make_dataset(n, nq, dim, metric=metric, rank=6, seed=7)
    graph = os.path.join(HERE, f"{name}.graph")
    rh.ref_slim_build(base, graph, metric=metric, M=M, ef_construction=60, branching="4", threads=1)
    ix = rh.RefSlim(graph, dim, n, metric, counting=True)
    out = {"queries": q, "n": n, "dim": dim, "metric": metric, "k": k}
    for ef in EFS:
        lab, per = ix.counts(q, k, ef)
        out[f"ref_labels_ef{ef}"] = lab
        out[f"ref_counts_ef{ef}"] = per
    gt, _ = rh.ref_bruteforce(base, q, 100, metric=metric)
    out["ref_gt100"] = gt
    out["ref_dist_samples"] = np.array([rh.ref_dist(q[i], base[i], metric) for i in range(nq)], np.float32)
    info = ix.info()
    out["info"] = np.array([info[x] for x in ("n", "maxlevel", "enterpoint", "maxM", "maxM0", "M")], np.int64)
    nodes = list(range(0, n, max(1, n // 40))) + [info["enterpoint"]]
    lv, lb, offs, ids = [], [], [0], []
    for i in nodes:
        for l in range(0, info["maxlevel"] + 1):
            level, label, nb = ix.node(i, l)
            if l == 0:
                lv.append(level)
                lb.append(label)
            ids.extend(nb.tolist() if l <= level else [])
            offs.append(len(ids))
    out.update(node_ids=np.array(nodes, np.uint32), node_level=np.array(lv, np.int32),
               node_label=np.array(lb, np.uint64), node_nbr_offsets=np.array(offs, np.int64),
               node_nbrs=np.array(ids, np.uint32))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, os.path.getsize(graph), "bytes graph; maxlevel", info["maxlevel"])


if __name__ == "__main__":
    make("slim_l2_2k", 2000, 16, 0, 8)
    make("slim_ip_1k", 1000, 16, 1, 8)
