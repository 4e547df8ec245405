"""CPU suite, part 2: host logic of the engine — loader/flattener, builder, file formats, and that
the C-ABI library loads and exports every symbol include/hnswslim_b200.h declares.
No compute entry point is called here (there is no GPU); they must fail loudly instead."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, HAVE_GPU, ROOT, get_corpus, needs_ref
from hnsw_slim_b200 import capi, vecs_io
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hnswslim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hs_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(capi.EXPORTS) <= declared
    assert L.hs_abi_version() == 1


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device error path")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(capi.HsError) as e:
        capi.Index(os.path.join(GOLDEN, "slim_l2_2k.graph"), 16)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.HsError) as e:
        capi.bruteforce_knn(np.zeros((4, 4), np.float32), np.zeros((1, 4), np.float32), 1)
    assert e.value.code == -3
    with pytest.raises(capi.HsError) as e:                     # the fused sharded exchange needs a device too
        capi.Exchange(0, 2, 0, 4, 100, 10)
    assert e.value.code == -3


def test_exchange_argument_errors():
    for bad in [(0, 0, 0, 4, 100, 10), (0, 2, 2, 4, 100, 10), (0, 17, 0, 4, 100, 10), (0, 2, 0, 0, 100, 10),
                (0, 2, 0, 4, 0, 10), (0, 2, 0, 4, 100, 0)]:
        with pytest.raises(capi.HsError) as e:
            capi.Exchange(*bad)
        assert e.value.code == -1, bad                          # HS_ERR_ARG before any CUDA call


def test_shardgroup_argument_errors():
    import ctypes as C
    L = capi.lib()
    fake = (C.c_void_p * 1)(1)                     # never dereferenced: the shape checks come first
    out = C.c_void_p()
    for world, rank, nq_max, k, depth in [(0, 0, 100, 10, 4), (2, 2, 100, 10, 4), (17, 0, 100, 10, 4),
                                          (2, 0, 0, 10, 4), (2, 0, 100, 0, 4), (2, 0, 100, 10, 1), (2, 0, 100, 10, 17)]:
        assert L.hs_shardgroup_create(fake, 1, world, rank, nq_max, k, depth, C.byref(out)) == -1
    assert L.hs_shardgroup_create(fake, 0, 1, 0, 100, 10, 4, C.byref(out)) == -1
    assert L.hs_shardgroup_create(None, 1, 1, 0, 100, 10, 4, C.byref(out)) == -1
    assert L.hs_shardgroup_submit(None, None, 1, None, None) == -1
    assert L.hs_shardgroup_wait(None) == -1 and L.hs_shardgroup_wait_oldest(None) == -1


@pytest.mark.parametrize("name,metric", [("slim_l2_2k", 0), ("slim_ip_1k", 1)])
def test_flattened_graph_matches_reference_layout(name, metric):
    graph = os.path.join(GOLDEN, name + ".graph")
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    dim = int(z["dim"])
    hg = capi.HostGraph(graph, dim)
    orc = rh.Oracle(graph, dim, metric)
    info, oinfo = hg.info(), orc.info()
    for key in ("n", "maxlevel", "threshold_level", "enterpoint", "maxM", "maxM0", "M", "ef_construction"):
        assert info[key] == oinfo[key], key
    assert info["dim_padded"] % 32 == 0 and info["deg0_stride"] % 32 == 0 and info["upper_stride"] % 8 == 0
    n_upper = sum_deg0 = max_deg0 = 0
    for i in range(info["n"]):
        lvl, lab, vec = hg.node(i)
        olvl, olab, ids0 = orc.node(i, 0)
        assert (lvl, lab) == (olvl, olab & 0xFFFFFFFF)
        assert np.array_equal(vec[:dim], orc.vector(i)) and not vec[dim:].any()
        for l in range(lvl + 1):
            assert np.array_equal(hg.row(i, l), orc.node(i, l)[2]), (i, l)
        assert len(hg.row(i, lvl + 1)) == 0
        n_upper += lvl > 0
        sum_deg0 += len(ids0)
        max_deg0 = max(max_deg0, len(ids0))
    assert (info["n_upper"], info["sum_deg0"], info["max_deg0"]) == (n_upper, sum_deg0, max_deg0)
    # golden accessors recorded from the reference itself
    offs, nbrs, c = z["node_nbr_offsets"], z["node_nbrs"], 0
    for node in z["node_ids"]:
        for l in range(info["maxlevel"] + 1):
            assert np.array_equal(hg.row(int(node), l), nbrs[offs[c]:offs[c + 1]])
            c += 1


def test_loader_rejects_bad_files(tmp_path):
    good = open(os.path.join(GOLDEN, "slim_l2_2k.graph"), "rb").read()
    cases = {"empty": b"", "header_only": good[:60], "truncated_records": good[:5000],
             "truncated_blobs": good[:-7], "garbage": bytes(range(256)) * 8}
    for name, data in cases.items():
        p = tmp_path / f"{name}.graph"
        p.write_bytes(data)
        with pytest.raises(capi.HsError) as e:
            capi.HostGraph(str(p), 16)
        assert e.value.code == -2, name
    with pytest.raises(capi.HsError) as e:                      # wrong dim
        capi.HostGraph(os.path.join(GOLDEN, "slim_l2_2k.graph"), 24)
    assert e.value.code == -2 and "dim" in str(e.value)
    with pytest.raises(capi.HsError) as e:
        capi.HostGraph(str(tmp_path / "nope.graph"), 16)
    assert e.value.code == -2 and "Cannot open" in str(e.value)


def _patched(data: bytes, off: int, fmt: str, value) -> bytes:
    import struct
    b = bytearray(data)
    struct.pack_into(fmt, b, off, value)
    return bytes(b)


def test_loader_rejects_corrupt_slimq_and_hnsw_headers(tmp_path):
    """Crafted headers must end in HS_ERR_IO, never in an out-of-bounds read or an exception across the C ABI.
    hnsw_slimq header (slimq.h:1161-1200): 93 bytes of the slim header, then num_cluster, dim, padded_dim,
    offset_cluster_id, offset_bin_data, offset_ex_data, size_bin_data, size_ex_data, ex_bits (u64 each)."""
    import struct
    good = open(os.path.join(GOLDEN, "slimq_d96.graph"), "rb").read()
    capi.HostGraph(os.path.join(GOLDEN, "slimq_d96.graph"), 96, capi.HS_KIND_SLIMQ).info()       # sanity: loads
    n, rec = struct.unpack_from("<QQ", good, 0)
    q0 = 93
    num_cluster, qdim, pd, off_c, off_b, off_e, size_bin, size_ex, ex_bits = struct.unpack_from("<9Q", good, q0)
    assert rec == 28 + size_bin + size_ex and off_e == 28 + size_bin
    cases = {
        "rec_too_small": _patched(good, 8, "<Q", 28 + 8),                    # record smaller than the code block
        "rec_huge": _patched(good, 8, "<Q", 1 << 62),                         # n * rec overflows
        "n_huge": _patched(good, 0, "<Q", (1 << 31) - 1),
        "size_ex_mismatch": _patched(good, q0 + 7 * 8, "<Q", size_ex + 8),
        "num_cluster_zero": _patched(good, q0, "<Q", 0),
        "num_cluster_huge": _patched(good, q0, "<Q", 1 << 40),
        "padded_dim_huge": _patched(good, q0 + 2 * 8, "<Q", 1 << 30),
        "maxM_huge": _patched(good, 60, "<Q", 1 << 20),
    }
    # a node whose cluster id points past the centroid table (the kernel would index g2c[] with it)
    elements_at = q0 + 9 * 8 + 1 + num_cluster * pd * 4 + 4 * pd // 8
    cases["cluster_id_out_of_range"] = _patched(good, elements_at + 24, "<I", num_cluster)
    for name, data in cases.items():
        pth = tmp_path / f"q_{name}.graph"
        pth.write_bytes(data)
        with pytest.raises(capi.HsError) as e:
            capi.HostGraph(str(pth), 96, capi.HS_KIND_SLIMQ)
        assert e.value.code == -2, name
    # plain HNSW files (hnsw.h:748-779): u64 offsetLevel0, max_elements, n, rec, label_offset, offsetData; i32 maxlevel;
    # u32 enterpoint; u64 maxM, maxM0, M; f64 mult; u64 efc
    hg = open(os.path.join(GOLDEN, "hnsw_l2_1k.graph"), "rb").read()
    hdim = (struct.unpack_from("<Q", hg, 24)[0] - 8 - 4 - 4 * struct.unpack_from("<Q", hg, 64)[0]) // 4
    capi.HostGraph(os.path.join(GOLDEN, "hnsw_l2_1k.graph"), hdim, capi.HS_KIND_HNSW).info()
    hcases = {
        "maxM0_wraps": _patched(hg, 64, "<Q", (1 << 62) + struct.unpack_from("<Q", hg, 64)[0]),   # 4 * maxM0 wraps to the same size
        "maxM_wraps": _patched(hg, 56, "<Q", (1 << 62) + struct.unpack_from("<Q", hg, 56)[0]),
        "rec_huge": _patched(hg, 24, "<Q", 1 << 62),
        "truncated": hg[: len(hg) // 2],
        "list_count_beyond_maxM0": _patched(hg, 100, "<H", 65535),
    }
    for name, data in hcases.items():
        pth = tmp_path / f"h_{name}.graph"
        pth.write_bytes(data)
        with pytest.raises(capi.HsError) as e:
            capi.HostGraph(str(pth), hdim, capi.HS_KIND_HNSW)
        assert e.value.code == -2, name
    # empty indices: maxlevel is -1 in the header and must not size anything
    empty = bytearray(good[:93 + 9 * 8 + 1 + num_cluster * pd * 4 + 4 * pd // 8])
    struct.pack_into("<Q", empty, 0, 0)
    struct.pack_into("<i", empty, 48, -1)
    pth = tmp_path / "q_empty.graph"
    pth.write_bytes(bytes(empty))
    assert capi.HostGraph(str(pth), 96, capi.HS_KIND_SLIMQ).info()["n"] == 0


def test_vecs_io_roundtrip(tmp_path):
    a = np.random.default_rng(1).standard_normal((17, 5)).astype(np.float32)
    b = np.arange(34, dtype=np.uint32).reshape(17, 2)
    vecs_io.write_vecs(str(tmp_path / "a.fvecs"), a)
    vecs_io.write_vecs(str(tmp_path / "b.ivecs"), b)
    assert np.array_equal(vecs_io.read_fvecs(str(tmp_path / "a.fvecs")), a)
    assert np.array_equal(vecs_io.read_ivecs(str(tmp_path / "b.ivecs")), b)
    raw = np.fromfile(str(tmp_path / "a.fvecs"), dtype=np.int32)
    assert raw[0] == 5 and raw.size == 17 * 6          # [int32 d][d floats] per row (util.h:149-168)


def test_synthetic_generator_is_prefix_stable():
    a, qa = make_dataset(1000, 10, 24, rank=6)
    b, qb = make_dataset(70000, 20, 24, rank=6)
    assert np.array_equal(a, b[:1000]) and np.array_equal(qa, qb[:10])
    c, _ = make_dataset(100, 1, 24, metric=1, rank=6)
    np.testing.assert_allclose(np.linalg.norm(c, axis=1), 1.0, rtol=1e-5)


def _graph_invariants(path, dim, n, M):
    orc = rh.Oracle(path, dim)
    info = orc.info()
    assert info["n"] == n and info["maxM0"] == 2 * M and info["maxM"] == M
    assert orc.node(info["enterpoint"], 0)[0] == info["maxlevel"]
    for i in range(0, n, max(1, n // 500)):
        lvl, label, _ = orc.node(i, 0)
        assert label == i
        for l in range(lvl + 1):
            ids = orc.node(i, l)[2]
            assert len(ids) <= (2 * M if l == 0 else M)
            assert len(set(ids.tolist())) == len(ids) and i not in ids and (ids < n).all()
            if l > 0:      # hierarchical pruning: upper-level neighbours live exactly on that level
                assert all(orc.node(int(j), 0)[0] == l for j in ids)


def test_builder_writes_the_reference_format(tmp_path):
    n, dim, M = 4000, 24, 8
    base, q = make_dataset(n, 100, dim, rank=6)
    path = str(tmp_path / "mine.graph")
    capi.build_slim_graph(base, path, M=M, ef_construction=80, threads=2)
    _graph_invariants(path, dim, n, M)
    orc = rh.Oracle(path, dim)
    assert np.array_equal(np.stack([orc.vector(i) for i in range(0, n, 97)]), base[::97])
    lab, _, _, _ = orc.search(q, 10, 80)
    gt, _ = rh.oracle_bruteforce(base, q, 10)
    rec = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(lab, gt)])
    assert rec > 0.9, rec
    # levels are a pure function of (seed, node): same seed => same levels whatever the thread timing
    capi.build_slim_graph(base, str(tmp_path / "again.graph"), M=M, ef_construction=80, threads=3)
    o2 = rh.Oracle(str(tmp_path / "again.graph"), dim)
    assert all(orc.node(i, 0)[0] == o2.node(i, 0)[0] for i in range(0, n, 13))
    # custom labels are stored as given
    labels = np.arange(n, dtype=np.uint64) + 10_000_000
    capi.build_slim_graph(base[:500], str(tmp_path / "lab.graph"), M=M, ef_construction=40, labels=labels[:500])
    o3 = rh.Oracle(str(tmp_path / "lab.graph"), dim)
    assert o3.node(7, 0)[1] == 10_000_007


@needs_ref
def test_reference_loads_and_searches_our_graph(tmp_path):
    """Drop-in check in the other direction: the reference's loadIndex + searchKnn on a .graph
    written by hs_build_slim_graph, with recall on par with a reference-built index."""
    c = get_corpus(n=20000, nq=300, dim=32)
    path = str(tmp_path / "mine.graph")
    capi.build_slim_graph(c.base, path, M=16, ef_construction=200)
    gt, _ = rh.ref_bruteforce(c.base, c.queries, 10)
    for ef in (32, 64):           # 0.5 pp, north_star's recall bar (measured spread over builder seeds: <= 0.3 pp)
        recalls = {}
        for name, g in (("ours", path), ("reference", c.graph)):
            ref = rh.RefSlim(g, c.dim, c.n)
            lab, _, _ = ref.search(c.queries, 10, ef)
            recalls[name] = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(lab, gt)])
        assert abs(recalls["ours"] - recalls["reference"]) <= 0.005, (ef, recalls)
    hi, ri = capi.HostGraph(path, c.dim).info(), capi.HostGraph(c.graph, c.dim).info()
    assert abs(hi["sum_deg0"] - ri["sum_deg0"]) / ri["sum_deg0"] < 0.05


def test_builder_argument_errors(tmp_path):
    base = np.zeros((10, 4), np.float32)
    with pytest.raises(capi.HsError):
        capi.build_slim_graph(base, str(tmp_path / "x.graph"), M=1)
    with pytest.raises(capi.HsError):
        capi.build_slim_graph(base, str(tmp_path / "x.graph"), branching="zero")
    with pytest.raises(capi.HsError):
        capi.build_slim_graph(base, "/nonexistent_dir/x.graph", M=4)


# ---------------------------------------------------------------- hnsw_slimq loader
@pytest.mark.parametrize("name", ["slimq_d96", "slimq_d128", "slimq_d200"])
def test_slimq_loader_matches_reference_accessors(name):
    """hs_debug_flatten on the reference-written hnsw_slimq .graph (slimq.h:1161-1216): header
    fields and level-0 rows against what the reference's own accessors returned (golden)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    graph = os.path.join(GOLDEN, f"{name}.graph")
    hg = capi.HostGraph(graph, int(g["dim"]), kind=capi.HS_KIND_SLIMQ)
    info = hg.info()
    ref = dict(zip(rh.RefSlimQ.INFO_KEYS, g["info"]))
    assert info["n"] == ref["n"] and info["maxlevel"] == ref["maxlevel"] and info["enterpoint"] == ref["enterpoint"]
    assert info["padded_dim_q"] == ref["padded_dim"] and info["num_cluster"] == ref["num_cluster"]
    assert info["kind"] == capi.HS_KIND_SLIMQ and info["M"] == ref["M"]
    offs = g["node_nbr_offsets"]
    for j, i in enumerate(g["node_ids"]):
        assert np.array_equal(hg.row(int(i), 0), g["node_nbrs"][offs[j]:offs[j + 1]])
    # a slim loader must refuse the slimq file and vice versa
    with pytest.raises(capi.HsError):
        capi.HostGraph(graph, int(g["dim"]), kind=capi.HS_KIND_SLIM)
    with pytest.raises(capi.HsError):
        capi.HostGraph(os.path.join(GOLDEN, "slim_l2_2k.graph"), 16, kind=capi.HS_KIND_SLIMQ)


@pytest.mark.skipif(rh.ref_slimq_path() is None, reason="needs oracle/_ref/libhsref_slimq_v4.so")
def test_reference_loads_engine_built_slimq_graph(tmp_path):
    """hs_build_slimq_graph writes HierarchicalNSWSlimQ::saveIndex's format (slimq.h:1161-1216): the
    reference loads it, searches it, and the restatement agrees with it id for id."""
    n, nq, dim, k = 8000, 100, 96, 10
    base, q = make_dataset(n, nq, dim, rank=10, seed=5)
    graph = str(tmp_path / "e.graph")
    capi.build_slimq_graph(base, graph, M=16, ef_construction=100, threads=4)
    r = rh.RefSlimQ(graph, base)
    assert r.info["n"] == n and r.info["padded_dim"] == 128 and r.info["num_cluster"] == 16
    o = rh.OracleQ(graph, base, t_const=r.t_const)
    gt, _ = rh.ref_bruteforce(base, q, k)
    rlab, _ = r.search(q, k, 100)
    olab, *_ = o.search(q, k, 100, order=rh.ORDER_REF)
    assert np.mean([set(a) == set(b) for a, b in zip(rlab, olab)]) >= 0.98
    rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(rlab, gt)])
    assert rec >= 0.95, rec
    # estimator quality of the engine's codes: estimates track the exact distances
    ids = np.arange(0, 2000, dtype=np.uint32)
    est = r.est(q[0], ids)
    exact = ((base[ids] - q[0]) ** 2).sum(1)
    assert np.corrcoef(est, exact)[0, 1] > 0.97
    # loader view
    hg = capi.HostGraph(graph, dim, kind=capi.HS_KIND_SLIMQ)
    for i in (0, 17, n - 1):
        assert np.array_equal(hg.row(i, 0), r.node(i)[3])


def test_every_environment_knob_is_documented_in_the_header():
    """The library reads a few tuning knobs from the environment (hs_load / plan time); each of them must be named
    in include/hnswslim_b200.h next to hs_set_tuning."""
    hdr = open(os.path.join(ROOT, "include", "hnswslim_b200.h")).read()
    names = set()
    csrc = os.path.join(ROOT, "hnsw_slim_b200", "csrc")
    for f in os.listdir(csrc):
        names |= set(re.findall(r'getenv\("([A-Z0-9_]+)"\)', open(os.path.join(csrc, f)).read()))
    names |= set(re.findall(r'environ\.get\("(HS_[A-Z0-9_]+)"', open(os.path.join(ROOT, "hnsw_slim_b200", "capi.py")).read()))
    assert len(names) >= 8
    missing = sorted(n for n in names if n not in hdr)
    assert not missing, f"environment knobs not documented in the header: {missing}"
