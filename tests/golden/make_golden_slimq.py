"""Generates the committed hnsw_slimq golden fixtures from the REFERENCE ITSELF
(oracle/_ref/libhsref_slimq_v4.so: the unmodified hnswalg_slimq.h + rabitqlib compiled by
oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_slimq.py

Outputs (small, committed), one pair per shape:
  slimq_d96.graph   2000 x 96  -> padded_dim 128, trunc_dim 64  (Kac-walk rotation path; DEEP/MSTuring shape)
  slimq_d128.graph  1000 x 128 -> padded_dim 128, trunc_dim 128 (pure FWHT rotation path; SIFT shape)
  slimq_d200.graph  1000 x 200 -> padded_dim 256, trunc_dim 128 (4 code words: generic kernel path)
  *.npz:
      queries, base_crc       the query batch; crc32 of the regenerated base rows
      t_const                 the query-quantiser constant the reference drew for this load
      ref_labels_ef{E}        the reference's searchKnn(q, k=10) labels per ef (heap order)
      ref_rotated/planes/scal/q2c   the reference's per-query preparation (slimq.h:1816-1847)
      ref_est_ids / ref_est   get_bin_est of query i against sampled nodes
      node_*                  cluster id / code words / factors / level-0 neighbours of sampled nodes
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EFS = (10, 40, 100)


def make(name, n, dim, M=8, nq=64, k=10):
    base, q = make_dataset(n, nq, dim, rank=6, seed=7)
    cent, cid = rh.kmeans(base, 16)
    graph = os.path.join(HERE, f"{name}.graph")
    rh.ref_slimq_build(base, cent, cid, graph, M=M, ef_construction=60, threads=1)
    ix = rh.RefSlimQ(graph, base)
    out = {"queries": q, "n": n, "dim": dim, "k": k, "base_crc": zlib.crc32(base.tobytes()),
           "t_const": ix.t_const}
    for ef in EFS:
        lab, _ = ix.search(q, k, ef)
        out[f"ref_labels_ef{ef}"] = lab
    rot, planes, scal, q2c = ix.prep(q)
    out.update(ref_rotated=rot, ref_planes=planes, ref_scal=scal, ref_q2c=q2c)
    ids = np.arange(0, n, max(1, n // 50), dtype=np.uint32)
    out["ref_est_ids"] = ids
    out["ref_est"] = np.stack([ix.est(q[i], ids) for i in range(nq)])
    out["info"] = np.array([ix.info[x] for x in rh.RefSlimQ.INFO_KEYS], np.int64)
    cl, codes, facs, offs, nbrs = [], [], [], [0], []
    for i in ids:
        c, code, fac, nb = ix.node(int(i))
        cl.append(c)
        codes.append(code)
        facs.append(fac)
        nbrs.extend(nb.tolist())
        offs.append(len(nbrs))
    out.update(node_ids=ids, node_cluster=np.array(cl, np.uint32), node_code=np.stack(codes),
               node_factors=np.stack(facs), node_nbr_offsets=np.array(offs, np.int64),
               node_nbrs=np.array(nbrs, np.uint32))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, os.path.getsize(graph), "bytes graph;", ix.info)


if __name__ == "__main__":
    make("slimq_d96", 2000, 96)
    make("slimq_d128", 1000, 128)
    make("slimq_d200", 1000, 200)
