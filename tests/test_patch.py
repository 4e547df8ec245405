"""Delta patches: hs_load_reserve + hs_patch_apply = the reference client's
loadIndex(path, space, max_elements) + patchFromStream (slim.h:2206-2388), fed by the streams the reference's
server side writes (convertFromHNSWWithDiff / genPatch, hnsw_slim_server_patch.cc:186-279).

CPU part: the host half of the patch path (parser, validation, upper-level re-slotting — the code the device
path runs, reached through hs_debug_patch) against the live reference client node by node and against the
committed fixture; the oracle's restatement (hso_patch) pinned to the reference's patched search.
GPU part: the HBM-resident index after patches == the oracle after the same patches, bit for bit; the saved
patched index == the reference client's saved file."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, HAVE_GPU, needs_ref
from hnsw_slim_b200 import capi
from hnsw_slim_b200.synth import make_dataset
from oracle import refharness as rh

K, EF = 10, 40


def _golden():
    z = np.load(os.path.join(GOLDEN, "patch_l2_1k.npz"))
    return z, os.path.join(GOLDEN, "patch_l2_1k.graph"), [z["patch0"].tobytes(), z["patch1"].tobytes()]


def _apply_golden(target, z, streams, upto=2):
    """patch0: rows from the caller's data (by label); patch1: rows inline."""
    infos = []
    for s in range(upto):
        if s == 0:
            infos.append(target.patch(streams[0], rows=z["base"]))
        else:
            infos.append(target.patch(streams[1], inline=True))
    return infos


def _check_nodes_against_golden(hg, z):
    n = int(z["n"])
    assert hg.info()["n"] == n
    offs, nbrs = z["node_nbr_offsets"], z["node_nbrs"]
    at = 0
    for i in range(n):
        lvl, lab, vec = hg.node(i)
        assert lvl == z["node_level"][i] and lab == z["node_label"][i], i
        assert np.array_equal(vec[: int(z["dim"])], z["base"][lab]), i
        for l in range(lvl + 1):
            assert np.array_equal(hg.row(i, l), nbrs[offs[at]:offs[at + 1]]), (i, l)
            at += 1
    assert at + 1 == len(offs)


def test_patch_golden_host_image():
    """Partial index + the reference's two patch streams -> exactly the reference client's patched index."""
    z, graph, streams = _golden()
    hg = capi.HostGraph(graph, int(z["dim"]))
    assert hg.info()["n"] == int(z["n0"])
    infos = _apply_golden(hg, z, streams)
    assert [i["n_after"] for i in infos] == [1050, 1200]
    assert all(i["bytes_consumed"] == len(s) for i, s in zip(infos, streams))
    assert all(i["changed_new"] == 150 for i in infos)
    _check_nodes_against_golden(hg, z)
    # entry point and maxlevel are the partial index's: patchFromStream does not touch them (slim.h:2206-2388)
    info = hg.info()
    assert (info["maxlevel"], info["enterpoint"]) == (int(z["info"][1]), int(z["info"][2]))


def test_patch_golden_oracle():
    """hso_patch (the restatement of patchFromStream) reproduces the reference client's answers at every stage."""
    z, graph, streams = _golden()
    q, dim = z["queries"], int(z["dim"])
    for stage in range(3):
        orc = rh.Oracle(graph, dim)
        for s in range(stage):
            orc.patch(streams[s], rows=None if s == 1 else z["base"], inline=(s == 1))
        lab, _, nd, _ = orc.search(q, K, EF, order=rh.ORDER_REF)
        want = z[f"ref_labels_s{stage}"]
        assert all(set(a) == set(b) for a, b in zip(lab, want)), stage
        assert np.array_equal(nd, z[f"ref_counts_s{stage}"]), stage


def test_patch_rows_by_label_subset():
    """patchFromStream(in, new_data) (slim.h:2343-2388): only the new labels' vectors are at hand."""
    z, graph, streams = _golden()
    hg = capi.HostGraph(graph, int(z["dim"]))
    sel = np.arange(890, 1060)[::-1].copy()                    # any order, a superset of the new labels
    hg.patch(streams[0], rows=z["base"][sel], row_labels=sel)
    hg.patch(streams[1], inline=True)
    _check_nodes_against_golden(hg, z)
    hg2 = capi.HostGraph(graph, int(z["dim"]))
    with pytest.raises(capi.HsError) as e:                     # a new node's vector is missing: nothing is applied
        hg2.patch(streams[0], rows=z["base"][:1000])
    assert e.value.code == -1 and "no vector for new node" in str(e.value)
    assert hg2.info()["n"] == int(z["n0"])


def _records(stream: bytes, dim: int, inline: bool):
    """Offsets of the records of a patch stream: [(offset of id, is_new, level, total, blob offset)]."""
    n_after, n_old, n_new = struct.unpack_from("<QQQ", stream, 0)
    pos, out = 24, []
    for i in range(n_old + n_new):
        new = i >= n_old
        nid, level, total = struct.unpack_from("<IiI", stream, pos)
        head = 12 + (8 if new else 0)
        bsz, = struct.unpack_from("<I", stream, pos + head)
        out.append((pos, new, level, total, pos + head + 4))
        pos += head + 4 + bsz + (4 * dim if new and inline else 0)
    assert pos == len(stream)
    return n_after, n_old, n_new, out


def test_patch_rejects_corrupt_streams():
    z, graph, streams = _golden()
    dim = int(z["dim"])
    good = streams[0]
    n_after, n_old, n_new, recs = _records(good, dim, False)
    assert (n_after, n_new) == (1050, 150)

    def expect(stream, code, text, **kw):
        hg = capi.HostGraph(graph, dim)
        with pytest.raises(capi.HsError) as e:
            hg.patch(stream, rows=z["base"], **kw)
        assert e.value.code == code and text in str(e.value), str(e.value)
        assert hg.info()["n"] == int(z["n0"])                 # unchanged

    expect(good[:20], -2, "truncated patch header")
    expect(good[: len(good) // 2], -2, "truncated")
    expect(good[:-3], -2, "truncated")
    bad = bytearray(good)
    struct.pack_into("<Q", bad, 8, 1 << 40)                    # changed_old_cnt
    expect(bytes(bad), -2, "record counts exceed")
    bad = bytearray(good)
    struct.pack_into("<Q", bad, 0, 10)                         # cur_element_count below the index's
    expect(bytes(bad), -2, "below the index")
    bad = bytearray(good)
    struct.pack_into("<I", bad, recs[0][0], 5000)              # node id
    expect(bytes(bad), -2, "out of range")
    bad = bytearray(good)
    struct.pack_into("<i", bad, recs[0][0] + 4, 77)            # level
    expect(bytes(bad), -2, "level")
    bad = bytearray(good)
    struct.pack_into("<I", bad, recs[0][0] + 8, recs[0][3] + 1)   # total no longer matches the blob size
    expect(bytes(bad), -2, "blob size mismatch")
    first_with_nbrs = next(r for r in recs if r[3] > 0)
    bad = bytearray(good)
    struct.pack_into("<I", bad, first_with_nbrs[4] + 2 * first_with_nbrs[2], 4000)   # first neighbour id
    expect(bytes(bad), -2, "neighbour id out of range")
    upper = next((r for r in recs if r[2] > 0 and r[3] > 0), None)
    if upper is not None:
        bad = bytearray(good)
        struct.pack_into("<H", bad, upper[4], 0xFFFF)          # offsets[0] beyond total
        expect(bytes(bad), -2, "corrupt level offsets")
    # an inline stream read as a plain one (and vice versa) falls apart at the first new record
    expect(streams[1], -2, "")
    hg = capi.HostGraph(graph, dim)
    with pytest.raises(capi.HsError):
        hg.patch(good, inline=True)


def test_patch_longer_lists_than_the_file_had():
    """hs_debug_flatten sizes rows for the longest list of the FILE; a patch that brings a longer one widens the
    host image (the device index is loaded with header-wide rows instead, hs_load_reserve)."""
    z, graph, _ = _golden()
    dim = int(z["dim"])
    hg = capi.HostGraph(graph, dim)
    before = hg.info()
    ids = np.arange(1, 41, dtype=np.uint32)                    # 40 level-0 neighbours for node 0 (stride is 32)
    stream = struct.pack("<QQQ", before["n"], 1, 0) + struct.pack("<IiI", 0, 0, 40) + struct.pack("<I", 160) + ids.tobytes()
    hg.patch(stream)
    assert hg.info()["deg0_stride"] == 64 and np.array_equal(hg.row(0, 0), ids)
    assert np.array_equal(hg.row(5, 0), capi.HostGraph(graph, dim).row(5, 0))


@needs_ref
@pytest.mark.parametrize("metric,dim,M", [(0, 32, 8), (1, 24, 16)])
def test_patch_host_image_matches_live_reference_client(tmp_path, metric, dim, M):
    """Three updates through the three patchFromStream overloads; every node of the patched host image equals the
    reference client's; the client's saved file flattens to the same image; oracle == reference on the result."""
    n, n0, nq = 5000, 3800, 150
    base, q = make_dataset(n, nq, dim, metric=metric, rank=8, seed=5)
    part, fin = str(tmp_path / "part.graph"), str(tmp_path / "final.graph")
    names = rh.ref_slim_make_patches(base, n0, 3, part, str(tmp_path / "p"), final_path=fin, inline_last=True,
                                     metric=metric, M=M, ef_construction=60, threads=1)
    streams = [open(nm, "rb").read() for nm in names]
    cli = rh.RefSlim(part, dim, n, metric)
    hg = capi.HostGraph(part, dim)
    orc = rh.Oracle(part, dim, metric)
    new_lo = [n0 + (n - n0) * r // 3 for r in range(4)]
    for r, st in enumerate(streams):
        if r == 0:                                             # the client's own path: a map of the new rows
            cli.patch(st, rh.PATCH_MAP, base)
            sel = np.arange(new_lo[0], new_lo[1])
            info = hg.patch(st, rows=base[sel], row_labels=sel)
            orc.patch(st, rows=base)
        elif r == 1:
            cli.patch(st, rh.PATCH_VECTORS, base)
            info = hg.patch(st, rows=base)
            orc.patch(st, rows=base)
        else:
            cli.patch(st, rh.PATCH_INLINE)
            info = hg.patch(st, inline=True)
            orc.patch(st, inline=True)
        assert info["n_after"] == new_lo[r + 1] == cli.info()["n"] and info["bytes_consumed"] == len(st)
    for i in range(n):
        lvl, lab, vec = hg.node(i)
        rl, rlab, _ = cli.node(i, 0)
        assert (lvl, lab) == (rl, rlab), i
        assert np.array_equal(vec[:dim], base[lab]), i
        for l in range(lvl + 1):
            assert np.array_equal(hg.row(i, l), cli.node(i, l)[2]), (i, l)
    saved = str(tmp_path / "client.graph")
    cli.save(saved)
    a, b = hg.info(), capi.HostGraph(saved, dim).info()
    assert {k: a[k] for k in a if k not in ("deg0_stride", "upper_stride")} == \
           {k: b[k] for k in b if k not in ("deg0_stride", "upper_stride")}
    # the oracle's restatement of patchFromStream == the reference's, by search result and evaluation count
    cc = rh.RefSlim(saved, dim, n, metric, counting=True)
    want, per = cc.counts(q, K, EF)
    got, _, nd, _ = orc.search(q, K, EF, order=rh.ORDER_REF)
    assert all(set(x) == set(y) for x, y in zip(got, want))
    assert np.array_equal(nd, per)
    # and the patched client answers like the server that produced the patches
    srv, _, _ = rh.RefSlim(fin, dim, n, metric).search(q, K, EF)
    same = np.mean([set(x) == set(y) for x, y in zip(want, srv)])
    assert same >= 0.97, same                                   # differs only through the entry point / maxlevel the client keeps


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device error path")
def test_patch_entry_points_need_a_device():
    _, graph, _ = _golden()
    with pytest.raises(capi.HsError) as e:
        capi.Index.load_reserve(graph, 16, 1200)
    assert e.value.code == -3
    import ctypes as C
    L, out = capi.lib(), C.c_void_p()
    assert L.hs_load_reserve(graph.encode(), capi.HS_KIND_SLIMQ, 0, 16, 1200, 0, C.byref(out)) == -5
    assert L.hs_load_reserve(graph.encode(), capi.HS_KIND_SLIM, 0, 16, 0, 0, C.byref(out)) == -1
    assert L.hs_patch_apply(None, None, 0, 0, None, None, 0, None) == -1


# ------------------------------------------------------------------------------------------------ GPU

def _assert_index_equals_oracle(ix, orc, q, ef=EF, k=K):
    ix.set_ef(ef)
    lab, dist, cnt = ix.search(q, k, counts=True)
    olab, odist, ond, onh = orc.search(q, k, ef, order=rh.ORDER_GPU, team=8)
    assert np.array_equal(lab, olab)
    assert np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    assert np.array_equal(cnt[:, 0], ond) and np.array_equal(cnt[:, 1], onh)


@pytest.mark.gpu
def test_gpu_patch_golden():
    """The HBM-resident index after each of the reference's patch streams == the oracle after the same streams
    (ids, distance bits, per-query counters), and its answers are the reference client's."""
    z, graph, streams = _golden()
    q, dim = z["queries"], int(z["dim"])
    ix = capi.Index.load_reserve(graph, dim, 1200)
    orc = rh.Oracle(graph, dim)
    assert ix.info()["deg0_stride"] == 32 and ix.info()["n"] == 900
    _assert_index_equals_oracle(ix, orc, q)
    for s in range(2):
        info = ix.patch(streams[s], rows=None if s == 1 else z["base"], inline=(s == 1))
        orc.patch(streams[s], rows=None if s == 1 else z["base"], inline=(s == 1))
        assert info["n_after"] == (1050, 1200)[s] == ix.info()["n"] and info["rows_written"] > 0
        _assert_index_equals_oracle(ix, orc, q)
        lab, _ = ix.search(q, K)
        want = z[f"ref_labels_s{s + 1}"]
        assert np.mean([set(a) == set(b) for a, b in zip(lab, want)]) >= 0.98
    for ef in (10, 100, 200):
        _assert_index_equals_oracle(ix, orc, q, ef=ef)


@pytest.mark.gpu
def test_gpu_patch_errors_leave_the_index_alone():
    z, graph, streams = _golden()
    q, dim = z["queries"], int(z["dim"])
    plain = capi.Index(graph, dim)
    with pytest.raises(capi.HsError) as e:
        plain.patch(streams[0], rows=z["base"])
    assert e.value.code == -5 and "hs_load_reserve" in str(e.value)
    tight = capi.Index.load_reserve(graph, dim, 1000)
    with pytest.raises(capi.HsError) as e:
        tight.patch(streams[0], rows=z["base"])                # 1050 > 1000
    assert e.value.code == -1 and "max_elements" in str(e.value)
    with pytest.raises(capi.HsError):
        capi.Index.load_reserve(graph, dim, 100)               # below the file's count
    ix = capi.Index.load_reserve(graph, dim, 1200)
    orc = rh.Oracle(graph, dim)
    for bad in (streams[0][:1000], streams[1]):
        with pytest.raises(capi.HsError):
            ix.patch(bad, rows=z["base"])
        assert ix.info()["n"] == 900
        _assert_index_equals_oracle(ix, orc, q)
    with pytest.raises(capi.HsError):
        ix.patch(streams[0], rows=z["base"][:950])             # a new node's vector is missing
    _assert_index_equals_oracle(ix, orc, q)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("metric,dim,M", [(0, 128, 16), (1, 96, 32), (0, 200, 8)])
def test_gpu_patch_matches_oracle_and_reference_client(tmp_path, metric, dim, M):
    """A stream of updates as the reference's server produces them, batches in flight around every patch:
    engine == oracle bit for bit after each one; the saved patched index is the reference client's saved file;
    recall follows the server's."""
    n, n0, nq = 20000, 15000, 300
    base, q = make_dataset(n, nq, dim, metric=metric, rank=10, seed=3)
    part, fin = str(tmp_path / "part.graph"), str(tmp_path / "final.graph")
    names = rh.ref_slim_make_patches(base, n0, 4, part, str(tmp_path / "p"), final_path=fin, inline_last=True,
                                     metric=metric, M=M, ef_construction=100, threads=1)
    streams = [open(nm, "rb").read() for nm in names]
    ix = capi.Index.load_reserve(part, dim, n, metric=metric)
    assert ix.info()["deg0_stride"] == 32 * ((2 * M + 31) // 32)
    orc = rh.Oracle(part, dim, metric)
    cli = rh.RefSlim(part, dim, n, metric)
    gt, _ = capi.bruteforce_knn(base, q, K, metric=metric)
    recalls = []
    for r, st in enumerate(streams):
        inline = r == len(streams) - 1
        ix.set_ef(EF)
        ix.search(q, K)                                        # a batch right before the patch
        info = ix.patch(st, rows=None if inline else base, inline=inline)
        orc.patch(st, rows=None if inline else base, inline=inline)
        cli.patch(st, rh.PATCH_INLINE if inline else rh.PATCH_MAP, None if inline else base)
        assert info["n_after"] == ix.info()["n"] == cli.info()["n"]
        for ef in (EF, 130):
            _assert_index_equals_oracle(ix, orc, q, ef=ef)
        lab, _ = ix.search(q, K)
        recalls.append(float(np.mean([len(set(a) & set(b)) / K for a, b in zip(lab, gt)])))
    assert recalls[-1] > recalls[0]                            # the rows that arrived are found
    ours, theirs = str(tmp_path / "ours.graph"), str(tmp_path / "client.graph")
    ix.save(ours)
    cli.save(theirs)
    a, b = np.fromfile(ours, np.uint8), np.fromfile(theirs, np.uint8)
    assert a.size == b.size
    # identical except the 8 stale pointer bytes of every element record (slim.h:740: the raw records are dumped)
    hdr = 6 * 8 + 3 * 4 + 4 * 8 + 1
    rec = 24 + 4 * dim
    keep = np.ones(a.size, bool)
    ptr = hdr + np.arange(n)[:, None] * rec + 16 + np.arange(8)[None, :]
    keep[ptr.ravel()] = False
    assert np.array_equal(a[keep], b[keep])
    srv, _, _ = rh.RefSlim(fin, dim, n, metric).search(q, K, EF)
    rec_srv = float(np.mean([len(set(x) & set(y)) / K for x, y in zip(srv, gt)]))
    ix.set_ef(EF)
    lab, _ = ix.search(q, K)
    rec_ours = float(np.mean([len(set(x) & set(y)) / K for x, y in zip(lab, gt)]))
    assert abs(rec_ours - rec_srv) <= 0.01, (rec_ours, rec_srv)


def test_patch_parser_survives_random_corruption():
    """Memory safety of the stream parser: 400 randomly corrupted copies of a valid stream (byte flips, truncations,
    duplicated slices) are each either rejected with a status code or applied to a consistent image — ids in range,
    rows padded, counts matching — and never crash the process."""
    z, graph, streams = _golden()
    dim = int(z["dim"])
    rng = np.random.default_rng(12345)
    good = np.frombuffer(streams[0], np.uint8)
    applied = rejected = 0
    for trial in range(400):
        b = good.copy()
        kind = trial % 4
        if kind == 0:                                            # a few byte flips anywhere
            at = rng.integers(0, b.size, size=rng.integers(1, 6))
            b[at] = rng.integers(0, 256, size=at.size, dtype=np.uint8)
        elif kind == 1:                                          # flips inside the header and the first records
            at = rng.integers(0, 200, size=rng.integers(1, 4))
            b[at] = rng.integers(0, 256, size=at.size, dtype=np.uint8)
        elif kind == 2:                                          # truncation
            b = b[: rng.integers(0, b.size)]
        else:                                                    # a slice duplicated in place
            lo = int(rng.integers(24, b.size - 64))
            ln = int(rng.integers(1, 64))
            b = np.concatenate([b[:lo], b[lo:lo + ln], b[lo:]])
        hg = capi.HostGraph(graph, dim)
        try:
            info = hg.patch(b.tobytes(), rows=z["base"])
        except capi.HsError as e:
            assert e.code in (-1, -2, -4, -5), e
            assert hg.info()["n"] == int(z["n0"])
            rejected += 1
            continue
        applied += 1
        n = hg.info()["n"]
        assert n == info["n_after"] and n >= int(z["n0"])      # a corrupted element count is a (legal) growth of the host image
        for i in rng.integers(0, n, size=40):
            lvl, _, _ = hg.node(int(i))
            for l in range(lvl + 1):
                r = hg.row(int(i), l)
                assert (r < n).all()
    assert rejected > 200 and applied + rejected == 400
