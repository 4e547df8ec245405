"""Generates the committed delta-patch fixture from the REFERENCE ITSELF (oracle/_ref): the server side of
hnsw_slim_server_patch.cc (addPoint + convertFromHNSWWithDiff / genPatch) writes the partial index and the
patch streams, the client side (loadIndex with max_elements + patchFromStream) applies them and is searched.
Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_patch.py

Outputs (small, committed):
  patch_l2_1k.graph       the PARTIAL index (rows [0, 900) of 1200 x 16 L2 vectors, M=8) as saveIndex wrote it
  patch_l2_1k.npz:
      base, queries
      patch0              /updateIndex response for rows [900, 1050): vectors come from the client's own data
      patch1              /getLastBatch form for rows [1050, 1200): vectors inline (genPatch, to_add = true)
      ref_labels_s{S}     the reference CLIENT's searchKnn(q, k=10, ef=40) labels after S patches (S = 0, 1, 2)
      ref_counts_s{S}     distance evaluations per query counted inside the reference's DISTFUNC
      node_*              level / label / neighbour slices of every node of the fully patched client index
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hnsw_slim_b200.synth import make_dataset  # noqa: E402
from oracle import refharness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N, N0, DIM, NQ, K, EF, M = 1200, 900, 16, 64, 10, 40, 8


def main():
    base, q = make_dataset(N, NQ, DIM, rank=6, seed=11)
    graph = os.path.join(HERE, "patch_l2_1k.graph")
    out = {"base": base, "queries": q, "n": N, "n0": N0, "dim": DIM, "k": K, "ef": EF}
    with tempfile.TemporaryDirectory() as td:
        names = rh.ref_slim_make_patches(base, N0, 2, graph, os.path.join(td, "p"), inline_last=True, M=M,
                                         ef_construction=60, threads=1)
        streams = [open(nm, "rb").read() for nm in names]
        # one client per stage: the counting DISTFUNC is chosen when the index is opened
        for stage in range(3):
            cli = rh.RefSlim(graph, DIM, N, counting=True)
            for s in range(stage):
                cli.patch(streams[s], rh.PATCH_INLINE if s == 1 else rh.PATCH_MAP, None if s == 1 else base)
            lab, per = cli.counts(q, K, EF)
            out[f"ref_labels_s{stage}"] = lab
            out[f"ref_counts_s{stage}"] = per
        info = cli.info()
        lv, lb, offs, ids = [], [], [0], []
        top = 0
        for i in range(N):
            level, label, _ = cli.node(i, 0)
            top = max(top, level)
            lv.append(level)
            lb.append(label)
            for l in range(level + 1):
                ids.extend(cli.node(i, l)[2].tolist())
                offs.append(len(ids))
        out.update(node_level=np.array(lv, np.int32), node_label=np.array(lb, np.uint64),
                   node_nbr_offsets=np.array(offs, np.int64), node_nbrs=np.array(ids, np.uint32),
                   info=np.array([info[x] for x in ("n", "maxlevel", "enterpoint", "maxM", "maxM0", "M")], np.int64))
    out["patch0"] = np.frombuffer(streams[0], np.uint8)
    out["patch1"] = np.frombuffer(streams[1], np.uint8)
    np.savez_compressed(os.path.join(HERE, "patch_l2_1k.npz"), **out)
    print("patch_l2_1k", os.path.getsize(graph), "bytes graph;", [len(s) for s in streams], "bytes patches; top level", top)


if __name__ == "__main__":
    main()
