"""TEST INFRASTRUCTURE — ctypes bindings for the checker libraries.

* ``RefSlim`` / ``ref_*``  -> oracle/_ref/libhsref_slim_{v3,v4}.so: the UNMODIFIED
  reference (hnswlib fork under /root/reference) compiled by oracle/Makefile.
* ``Oracle``               -> oracle/_build/libhs_oracle.so: our plain-C restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (hnsw_slim_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "_build", "libhs_oracle.so")
REFERENCE_ROOT = "/root/reference"

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def cpu_level() -> str:
    """'v4' if the host CPU runs x86-64-v4 (AVX-512 F/BW/CD/DQ/VL) code, else 'v3'."""
    fl = _cpu_flags()
    need = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"}
    return "v4" if need <= fl else "v3"


def have_reference_tree() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "third_party", "hnswlib"))


def build(ref: bool = True, oracle: bool = True, quiet: bool = True) -> None:
    """Compile the checker libraries (make).  `ref` needs /root/reference."""
    targets = []
    if oracle:
        targets.append("oracle")
    if ref and have_reference_tree():
        targets.append("ref")
    if not targets:
        return
    subprocess.run(["make", "-C", HERE, "-j4"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def ref_slim_path() -> str | None:
    p = os.path.join(REF_DIR, f"libhsref_slim_{cpu_level()}.so")
    return p if os.path.exists(p) else None


def ref_slimq_path() -> str | None:
    if cpu_level() != "v4" or "avx512_vpopcntdq" not in _cpu_flags():
        return None
    p = os.path.join(REF_DIR, "libhsref_slimq_v4.so")
    return p if os.path.exists(p) else None


_slim_lib = None


def slim_lib():
    global _slim_lib
    if _slim_lib is None:
        p = ref_slim_path()
        if p is None:
            raise RuntimeError("oracle/_ref/libhsref_slim_*.so not built (run `make -C oracle ref` "
                               "where /root/reference exists)")
        L = C.CDLL(p)
        L.ref_last_error.restype = C.c_char_p
        L.ref_num_procs.restype = C.c_int
        L.ref_slim_build.restype = C.c_int
        L.ref_slim_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t,
                                     C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_slim_open.restype = C.c_void_p
        L.ref_slim_open.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t, C.c_int]
        L.ref_slim_close.argtypes = [C.c_void_p]
        L.ref_slim_info.argtypes = [C.c_void_p, _u64p]
        L.ref_slim_node.restype = C.c_int
        L.ref_slim_node.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_uint64), _u32p, C.c_int]
        L.ref_slim_search.restype = C.c_int
        L.ref_slim_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                      _u32p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.ref_slim_counts.restype = C.c_int
        L.ref_slim_counts.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p, _u64p]
        L.ref_dist.restype = C.c_float
        L.ref_dist.argtypes = [_f32p, _f32p, C.c_size_t, C.c_int]
        if hasattr(L, "ref_strategy_recall"):
            L.ref_strategy_recall.restype = C.c_double
            L.ref_strategy_recall.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.ref_bruteforce.restype = C.c_int
        L.ref_bruteforce.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, _f32p, C.c_size_t, C.c_size_t,
                                     C.c_int, _u32p, C.c_void_p, C.POINTER(C.c_double)]
        if hasattr(L, "ref_slim_make_patches"):
            L.ref_slim_make_patches.restype = C.c_int
            L.ref_slim_make_patches.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t,
                                                C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_size_t, C.c_size_t,
                                                C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                                C.c_char_p, C.c_char_p, C.c_char_p]
            L.ref_slim_patch.restype = C.c_int
            L.ref_slim_patch.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
            L.ref_slim_save.restype = C.c_int
            L.ref_slim_save.argtypes = [C.c_void_p, C.c_char_p]
        _slim_lib = L
    return _slim_lib


def ref_hnsw_path() -> str | None:
    p = os.path.join(REF_DIR, f"libhsref_hnsw_{cpu_level()}.so")
    return p if os.path.exists(p) else None


_hnsw_lib = None


def hnsw_lib():
    """oracle/_ref/libhsref_hnsw_*.so: the reference's plain HierarchicalNSW and HierarchicalNSWSlimZero."""
    global _hnsw_lib
    if _hnsw_lib is None:
        p = ref_hnsw_path()
        if p is None:
            raise RuntimeError("oracle/_ref/libhsref_hnsw_*.so not built (run `make -C oracle ref`)")
        L = C.CDLL(p)
        L.refh_last_error.restype = C.c_char_p
        L.refh_hnsw_build.restype = C.c_int
        L.refh_hnsw_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_char_p,
                                      C.c_int, C.c_void_p, C.c_char_p]
        L.refh_hnsw_open.restype = C.c_void_p
        L.refh_hnsw_open.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t]
        L.refh_hnsw_close.argtypes = [C.c_void_p]
        L.refh_hnsw_search.restype = C.c_int
        L.refh_hnsw_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _u32p,
                                       C.c_void_p, C.POINTER(C.c_double)]
        L.refh_slimzero_build.restype = C.c_int
        L.refh_slimzero_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_char_p,
                                          C.c_int, C.c_float, C.c_float, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_char_p]
        L.refh_slimzero_open.restype = C.c_void_p
        L.refh_slimzero_open.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t]
        L.refh_slimzero_close.argtypes = [C.c_void_p]
        L.refh_slimzero_search.restype = C.c_int
        L.refh_slimzero_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p,
                                           C.POINTER(C.c_double)]
        _hnsw_lib = L
    return _hnsw_lib


def _run_isolated(mode: str, base: np.ndarray, args: dict) -> None:
    """Reference builders run in a child process (thread_local scratch in convertFromHNSW, see ref_slim_build)."""
    import json
    import sys
    import tempfile
    with tempfile.TemporaryDirectory(prefix="hsref_") as td:
        np.save(os.path.join(td, "base.npy"), np.ascontiguousarray(base, dtype=np.float32))
        with open(os.path.join(td, "args.json"), "w") as f:
            json.dump(args, f)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), mode, td], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference builder failed (rc={r.returncode}): {r.stderr[-2000:]}")


def ref_hnsw_build(base: np.ndarray, path: str, *, metric: int = 0, M: int = 16, ef_construction: int = 200,
                   branching: str = "4", threads: int = 0, isolate: bool = True) -> None:
    """hnsw_strategy.h:24-45: omp addPoint + HierarchicalNSW::saveIndex."""
    if isolate:
        return _run_isolated("--build-hnsw", base, dict(path=path, metric=metric, M=M, ef_construction=ef_construction,
                                                        branching=branching, threads=threads))
    L = hnsw_lib()
    base = np.ascontiguousarray(base, dtype=np.float32)
    if L.refh_hnsw_build(base, base.shape[0], base.shape[1], metric, M, ef_construction, branching.encode(), threads,
                         None, path.encode()) != 0:
        raise RuntimeError(L.refh_last_error().decode())


def ref_slimzero_build(base: np.ndarray, path: str, *, metric: int = 0, M: int = 16, ef_construction: int = 200,
                       branching: str = "4", threads: int = 0, min_indegree0: int = 8, min_indegree: int = 4,
                       isolate: bool = True, **prune) -> None:
    """hnsw_slimzero_strategy.h:38-103: HNSW build + HierarchicalNSWSlimZero::convertFromHNSW + saveIndex."""
    if isolate:
        return _run_isolated("--build-slimzero", base,
                             dict(path=path, metric=metric, M=M, ef_construction=ef_construction, branching=branching,
                                  threads=threads, min_indegree0=min_indegree0, min_indegree=min_indegree, prune=prune))
    L = hnsw_lib()
    p = dict(PRUNE_DEFAULTS)
    p.update(prune)
    base = np.ascontiguousarray(base, dtype=np.float32)
    rc = L.refh_slimzero_build(base, base.shape[0], base.shape[1], metric, M, ef_construction, branching.encode(),
                               p["threshold_level"], p["top_degree_percent0"], p["top_degree_percent"], p["top_M0"],
                               p["low_m0"], p["top_M"], p["low_m"], min_indegree0, min_indegree, threads, None,
                               path.encode())
    if rc != 0:
        raise RuntimeError(L.refh_last_error().decode())


class RefHnsw:
    """The reference's plain HierarchicalNSW<float> (the `hnsw` strategy) loaded from its .graph file."""

    def __init__(self, path: str, dim: int, n: int, metric: int = 0):
        self.L = hnsw_lib()
        self.h = self.L.refh_hnsw_open(path.encode(), dim, metric, n)
        if not self.h:
            raise RuntimeError(self.L.refh_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refh_hnsw_close(self.h)
            self.h = None

    def search(self, q: np.ndarray, k: int, ef: int, threads: int = 1):
        """-> labels[nq,k] nearest first, dists[nq,k], seconds."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        lab = np.zeros((q.shape[0], k), dtype=np.uint32)
        dist = np.zeros((q.shape[0], k), dtype=np.float32)
        sec = C.c_double(0)
        self.L.refh_hnsw_search(self.h, q, q.shape[0], k, ef, threads, lab, dist.ctypes.data, C.byref(sec))
        return lab, dist, sec.value


class RefSlimZero:
    """The reference's HierarchicalNSWSlimZero<float> loaded from its .graph file."""

    def __init__(self, path: str, dim: int, n: int, metric: int = 0):
        self.L = hnsw_lib()
        self.h = self.L.refh_slimzero_open(path.encode(), dim, metric, n)
        if not self.h:
            raise RuntimeError(self.L.refh_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refh_slimzero_close(self.h)
            self.h = None

    def search(self, q: np.ndarray, k: int, ef: int):
        """-> labels[nq,k] (unordered within a row), seconds."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        lab = np.zeros((q.shape[0], k), dtype=np.uint32)
        sec = C.c_double(0)
        self.L.refh_slimzero_search(self.h, q, q.shape[0], k, ef, lab, C.byref(sec))
        return lab, sec.value


# pruning parameters = main.cc defaults (main.cc:27-35,58-70)
PRUNE_DEFAULTS = dict(threshold_level=0, top_degree_percent0=0.02, top_degree_percent=0.02,
                      top_M0=32, low_m0=8, top_M=16, low_m=4)


def ref_slim_build(base: np.ndarray, path: str, *, metric: int = 0, M: int = 16, ef_construction: int = 200,
                   branching: str = "4", threads: int = 0, labels: np.ndarray | None = None,
                   hnsw_path: str = "", isolate: bool = True, **prune) -> tuple[float, float]:
    """Reference builder: omp addPoint -> convertFromHNSW -> saveIndex (hnsw_slim_strategy.h:60-95).

    isolate=True runs it in a fresh process: convertFromHNSW keeps `thread_local` scratch sized
    from the FIRST index a thread converts (slim.h:957-959,1008-1011), so a second build with a
    larger M in the same process overruns it.  The reference never builds twice per process.
    """
    if isolate:
        import json
        import sys
        import tempfile
        with tempfile.TemporaryDirectory(prefix="hsref_") as td:
            np.save(os.path.join(td, "base.npy"), np.ascontiguousarray(base, dtype=np.float32))
            if labels is not None:
                np.save(os.path.join(td, "labels.npy"), np.ascontiguousarray(labels, dtype=np.uint64))
            args = dict(path=path, metric=metric, M=M, ef_construction=ef_construction, branching=branching,
                        threads=threads, hnsw_path=hnsw_path, prune=prune, have_labels=labels is not None)
            with open(os.path.join(td, "args.json"), "w") as f:
                json.dump(args, f)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--build-slim", td],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"reference builder failed (rc={r.returncode}): {r.stderr[-2000:]}")
            bs, cs = json.loads(r.stdout.strip().splitlines()[-1])
            return bs, cs
    L = slim_lib()
    p = dict(PRUNE_DEFAULTS)
    p.update(prune)
    base = np.ascontiguousarray(base, dtype=np.float32)
    n, dim = base.shape
    bs, cs = C.c_double(0), C.c_double(0)
    lab = None
    if labels is not None:
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        lab = labels.ctypes.data_as(C.c_void_p)
    rc = L.ref_slim_build(base, n, dim, metric, M, ef_construction, branching.encode(),
                          p["threshold_level"], p["top_degree_percent0"], p["top_degree_percent"],
                          p["top_M0"], p["low_m0"], p["top_M"], p["low_m"], threads, lab,
                          path.encode(), hnsw_path.encode(), C.byref(bs), C.byref(cs))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    return bs.value, cs.value


PATCH_MAP, PATCH_VECTORS, PATCH_INLINE = 0, 1, 2   # the three patchFromStream overloads (ref_slim_patch)


def ref_slim_make_patches(base: np.ndarray, n0: int, rounds: int, partial_path: str, patch_prefix: str, *,
                          final_path: str = "", inline_last: bool = False, metric: int = 0, M: int = 16,
                          ef_construction: int = 200, branching: str = "4", threads: int = 1,
                          isolate: bool = True, **prune) -> list[str]:
    """The reference's SERVER side of the delta-patch protocol (hnsw_slim_server_patch.cc:186-279): the index
    over rows [0, n0) -> partial_path, then `rounds` updates (addPoint + convertFromHNSWWithDiff), one patch
    stream per round -> patch_prefix + "<r>.bin".  Returns the patch file names."""
    names = [f"{patch_prefix}{r}.bin" for r in range(rounds)]
    if isolate:
        _run_isolated("--make-patches", base, dict(n0=n0, rounds=rounds, partial_path=partial_path,
                                                   patch_prefix=patch_prefix, final_path=final_path,
                                                   inline_last=inline_last, metric=metric, M=M,
                                                   ef_construction=ef_construction, branching=branching,
                                                   threads=threads, prune=prune))
        return names
    L = slim_lib()
    p = dict(PRUNE_DEFAULTS)
    p.update(prune)
    base = np.ascontiguousarray(base, dtype=np.float32)
    n, dim = base.shape
    rc = L.ref_slim_make_patches(base, n, dim, metric, M, ef_construction, branching.encode(), p["threshold_level"],
                                 p["top_degree_percent0"], p["top_degree_percent"], p["top_M0"], p["low_m0"],
                                 p["top_M"], p["low_m"], threads, n0, rounds, int(inline_last),
                                 partial_path.encode(), patch_prefix.encode(), final_path.encode())
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    return names


class RefSlim:
    """The reference's HierarchicalNSWSlim<float> loaded from a .graph file."""

    def __init__(self, path: str, dim: int, n: int, metric: int = 0, counting: bool = False):
        self.L = slim_lib()
        self.h = self.L.ref_slim_open(path.encode(), dim, metric, n, int(counting))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.dim = dim

    def close(self):
        if self.h:
            self.L.ref_slim_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def info(self) -> dict:
        a = np.zeros(10, dtype=np.uint64)
        self.L.ref_slim_info(self.h, a)
        keys = ["n", "size_data_per_element", "maxlevel", "threshold_level", "enterpoint", "maxM", "maxM0",
                "M", "ef_construction", "has_deleted"]
        d = {k: int(v) for k, v in zip(keys, a)}
        for k in ("maxlevel", "threshold_level"):
            if d[k] >= 1 << 63:
                d[k] -= 1 << 64
        return d

    def node(self, i: int, level: int):
        out = np.zeros(256, dtype=np.uint32)
        lvl, lab = C.c_int(0), C.c_uint64(0)
        cnt = self.L.ref_slim_node(self.h, i, level, C.byref(lvl), C.byref(lab), out, 256)
        return lvl.value, lab.value, out[:cnt].copy()

    def search(self, q: np.ndarray, k: int, ef: int, threads: int = 1):
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((q.shape[0], k), dtype=np.uint32)
        sec, calls = C.c_double(0), C.c_uint64(0)
        self.L.ref_slim_search(self.h, q, q.shape[0], k, ef, threads, out, C.byref(sec), C.byref(calls))
        return out, sec.value, calls.value

    def counts(self, q: np.ndarray, k: int, ef: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((q.shape[0], k), dtype=np.uint32)
        per = np.zeros(q.shape[0], dtype=np.uint64)
        self.L.ref_slim_counts(self.h, q, q.shape[0], k, ef, out, per)
        return out, per

    def patch(self, stream: bytes, mode: int = PATCH_MAP, rows: np.ndarray | None = None) -> None:
        """patchFromStream (slim.h:2206-2388); rows[label] = the vector of external label `label` (modes 0, 1)."""
        ptr, n_rows = None, 0
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            ptr, n_rows = rows.ctypes.data_as(C.c_void_p), rows.shape[0]
        if self.L.ref_slim_patch(self.h, stream, len(stream), mode, ptr, n_rows) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def save(self, path: str) -> None:
        if self.L.ref_slim_save(self.h, path.encode()) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())


def ref_strategy_recall(base, q, knn, gt, K: int) -> float:
    """The reference's OWN SolveStrategy::recall (solve_strategy.h:67-103) executed on temp .fvecs/.ivecs files
    (it prints with 6 significant digits)."""
    import tempfile
    from hnsw_slim_b200 import vecs_io
    with tempfile.TemporaryDirectory() as td:
        paths = [os.path.join(td, n) for n in ("base.fvecs", "query.fvecs", "knn.ivecs", "gt.ivecs")]
        vecs_io.write_vecs(paths[0], np.ascontiguousarray(base, dtype=np.float32))
        vecs_io.write_vecs(paths[1], np.ascontiguousarray(q, dtype=np.float32))
        vecs_io.write_vecs(paths[2], np.ascontiguousarray(knn, dtype=np.uint32))
        vecs_io.write_vecs(paths[3], np.ascontiguousarray(gt, dtype=np.uint32))
        r = slim_lib().ref_strategy_recall(*[p.encode() for p in paths], K)
    if r < 0:
        raise RuntimeError(slim_lib().ref_last_error().decode())
    return float(r)


def ref_dist(a, b, metric=0) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(slim_lib().ref_dist(a, b, a.shape[0], metric))


def ref_bruteforce(base, q, k, metric=0, threads=0, want_dists=False):
    """BruteForce::solve rows: k labels per query, FARTHEST first (brute_force_strategy.h:24-36)."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.zeros((q.shape[0], k), dtype=np.uint32)
    d = np.zeros((q.shape[0], k), dtype=np.float32) if want_dists else None
    sec = C.c_double(0)
    rc = slim_lib().ref_bruteforce(base, base.shape[0], base.shape[1], metric, q, q.shape[0], k, threads, out,
                                   d.ctypes.data_as(C.c_void_p) if want_dists else None, C.byref(sec))
    if rc != 0:
        raise RuntimeError(slim_lib().ref_last_error().decode())
    return (out, d, sec.value) if want_dists else (out, sec.value)


# --------------------------------------------------------------------------- hnsw_slimq reference
_slimq_lib = None


def slimq_lib():
    global _slimq_lib
    if _slimq_lib is None:
        p = ref_slimq_path()
        if p is None:
            raise RuntimeError("oracle/_ref/libhsref_slimq_v4.so not built or host CPU lacks AVX-512 VPOPCNTDQ")
        L = C.CDLL(p)
        L.refq_last_error.restype = C.c_char_p
        L.refq_build.restype = C.c_int
        L.refq_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u32p, C.c_size_t, C.c_size_t,
                                 C.c_int, C.c_float, C.c_float, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                 C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.refq_open.restype = C.c_void_p
        L.refq_open.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, _f32p]
        L.refq_close.argtypes = [C.c_void_p]
        L.refq_info.argtypes = [C.c_void_p, _u64p]
        L.refq_get_tconst.restype = C.c_double
        L.refq_get_tconst.argtypes = [C.c_void_p]
        L.refq_set_tconst.argtypes = [C.c_void_p, C.c_double]
        L.refq_search.restype = C.c_int
        L.refq_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p,
                                  C.POINTER(C.c_double)]
        L.refq_prep.restype = C.c_int
        L.refq_prep.argtypes = [C.c_void_p, _f32p, _f32p, _u64p, _f32p, _f32p]
        L.refq_est.restype = C.c_int
        L.refq_est.argtypes = [C.c_void_p, _f32p, _u32p, C.c_size_t, _f32p]
        L.refq_rotate.restype = C.c_int
        L.refq_rotate.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p]
        L.refq_node.restype = C.c_int
        L.refq_node.argtypes = [C.c_void_p, C.c_uint32, _u32p, _u64p, _f32p, _u32p, C.c_int]
        _slimq_lib = L
    return _slimq_lib


def kmeans(base: np.ndarray, k: int = 16, iters: int = 10, seed: int = 7, sample: int = 100000):
    """Lloyd k-means -> (centroids[k,dim] f32, cluster_ids[n] u32).  The reference has no producer for
    its *_centroids_16.fvecs / *_clusterids_16.ivecs inputs (hnsw_slimq_strategy.h:42-45)."""
    rng = np.random.default_rng(seed)
    base = np.ascontiguousarray(base, dtype=np.float32)
    n = base.shape[0]
    train = base[rng.choice(n, size=min(n, sample), replace=False)]
    cent = train[rng.choice(train.shape[0], size=k, replace=False)].copy()

    def assign(x):
        d = (x * x).sum(1)[:, None] - 2.0 * x @ cent.T + (cent * cent).sum(1)[None, :]
        return d.argmin(1)
    for _ in range(iters):
        a = assign(train)
        for c in range(k):
            m = a == c
            if m.any():
                cent[c] = train[m].mean(0)
    ids = np.concatenate([assign(base[i:i + 262144]) for i in range(0, n, 262144)]).astype(np.uint32)
    return cent.astype(np.float32), ids


def ref_slimq_build(base, centroids, cluster_ids, path: str, *, M: int = 32, ef_construction: int = 128,
                    threads: int = 1, isolate: bool = True, **prune) -> tuple[float, float]:
    """rabitqlib HNSW construct -> HierarchicalNSWSlimQ::convertFromHNSW -> saveIndex
    (hnsw_slimq_strategy.h:100-142).  threads=1 keeps internal id == label."""
    if isolate:
        import json
        import sys
        import tempfile
        with tempfile.TemporaryDirectory(prefix="hsrefq_") as td:
            np.save(os.path.join(td, "base.npy"), np.ascontiguousarray(base, dtype=np.float32))
            np.save(os.path.join(td, "cent.npy"), np.ascontiguousarray(centroids, dtype=np.float32))
            np.save(os.path.join(td, "cid.npy"), np.ascontiguousarray(cluster_ids, dtype=np.uint32))
            with open(os.path.join(td, "args.json"), "w") as f:
                json.dump(dict(path=path, M=M, ef_construction=ef_construction, threads=threads, prune=prune), f)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--build-slimq", td],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"reference slimq builder failed (rc={r.returncode}): {r.stderr[-2000:]}")
            return tuple(json.loads(r.stdout.strip().splitlines()[-1]))
    L = slimq_lib()
    p = dict(PRUNE_DEFAULTS)
    p.update(prune)
    base = np.ascontiguousarray(base, dtype=np.float32)
    centroids = np.ascontiguousarray(centroids, dtype=np.float32)
    cluster_ids = np.ascontiguousarray(cluster_ids, dtype=np.uint32)
    n, dim = base.shape
    bs, cs = C.c_double(0), C.c_double(0)
    rc = L.refq_build(base, n, dim, centroids, centroids.shape[0], cluster_ids, M, ef_construction,
                      p["threshold_level"], p["top_degree_percent0"], p["top_degree_percent"], p["top_M0"],
                      p["low_m0"], p["top_M"], p["low_m"], threads, path.encode(), C.byref(bs), C.byref(cs))
    if rc != 0:
        raise RuntimeError(L.refq_last_error().decode())
    return bs.value, cs.value


def ref_slimq_search_copies(path: str, base, q, k: int, ef: int, threads: int = 0, t_const: float = 0.0,
                            passes: int = 1):
    """All-core hnsw_slimq baseline: one index copy per OpenMP thread over one shared dataset (the reference's
    slimq search is not re-entrant).  -> labels[nq,k], seconds of the median pass, threads used."""
    L = slimq_lib()
    if not hasattr(L, "refq_search_copies"):
        raise RuntimeError("oracle/_ref/libhsref_slimq was built before refq_search_copies existed")
    L.refq_search_copies.restype = C.c_int
    L.refq_search_copies.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, _f32p, C.c_int, _f32p, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.c_double, C.c_int, _u32p, C.POINTER(C.c_double)]
    base = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    threads = threads or (os.cpu_count() or 1)
    out = np.zeros((q.shape[0], k), dtype=np.uint32)
    sec = C.c_double(0)
    if L.refq_search_copies(path.encode(), base.shape[1], base.shape[0], base, threads, q, q.shape[0], k, ef,
                            float(t_const), passes, out, C.byref(sec)) != 0:
        raise RuntimeError(L.refq_last_error().decode())
    return out, sec.value, threads


class RefSlimQ:
    """The reference's HierarchicalNSWSlimQ<float> loaded from a .graph file + setDataset(base)."""

    INFO_KEYS = ["n", "size_data_per_element", "maxM", "maxM0", "M", "ef_construction", "maxlevel",
                 "threshold_level", "enterpoint", "num_cluster", "dim", "padded_dim", "ex_bits", "metric_type"]

    def __init__(self, path: str, base: np.ndarray):
        self.L = slimq_lib()
        base = np.ascontiguousarray(base, dtype=np.float32)
        self.dim = base.shape[1]
        self.h = self.L.refq_open(path.encode(), self.dim, base.shape[0], base)
        if not self.h:
            raise RuntimeError(self.L.refq_last_error().decode())
        a = np.zeros(len(self.INFO_KEYS), dtype=np.uint64)
        self.L.refq_info(self.h, a)
        self.info = {k: int(v) for k, v in zip(self.INFO_KEYS, a)}
        for k in ("maxlevel", "threshold_level"):
            if self.info[k] >= 1 << 63:
                self.info[k] -= 1 << 64

    def close(self):
        if self.h:
            self.L.refq_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def t_const(self) -> float:
        return float(self.L.refq_get_tconst(self.h))

    @t_const.setter
    def t_const(self, v: float):
        self.L.refq_set_tconst(self.h, float(v))

    def search(self, q, k: int, ef: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((q.shape[0], k), dtype=np.uint32)
        sec = C.c_double(0)
        if self.L.refq_search(self.h, q, q.shape[0], k, ef, out, C.byref(sec)) != 0:
            raise RuntimeError(self.L.refq_last_error().decode())
        return out, sec.value

    def prep(self, q):
        """-> rotated[nq,pd], planes[nq,pd/64*4] u64, scal[nq,3] (delta, vl, k1xsumq), q2c[nq,ncl]"""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        pd, nc = self.info["padded_dim"], self.info["num_cluster"]
        nq = q.shape[0]
        rot = np.zeros((nq, pd), np.float32)
        planes = np.zeros((nq, pd // 64 * 4), np.uint64)
        scal = np.zeros((nq, 3), np.float32)
        q2c = np.zeros((nq, nc), np.float32)
        for i in range(nq):
            self.L.refq_prep(self.h, q[i], rot[i], planes[i], scal[i], q2c[i])
        return rot, planes, scal, q2c

    def est(self, q, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.zeros(ids.shape[0], np.float32)
        self.L.refq_est(self.h, np.ascontiguousarray(q, dtype=np.float32), ids, ids.shape[0], out)
        return out

    def rotate(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, self.dim)
        out = np.zeros((v.shape[0], self.info["padded_dim"]), np.float32)
        self.L.refq_rotate(self.h, v, v.shape[0], out)
        return out

    def node(self, i: int):
        """-> cluster id, code words u64[pd/64], (f_add, f_rescale, f_error), level-0 neighbour ids"""
        cl = np.zeros(1, np.uint32)
        code = np.zeros(self.info["padded_dim"] // 64, np.uint64)
        fac = np.zeros(3, np.float32)
        nb = np.zeros(256, np.uint32)
        cnt = self.L.refq_node(self.h, i, cl, code, fac, nb, 256)
        return int(cl[0]), code, fac, nb[:cnt].copy()


# --------------------------------------------------------------------------- C restatement
ORDER_SEQ, ORDER_REF, ORDER_GPU, ORDER_SEQFMA = 0, 1, 2, 3
_oracle_lib = None


class _HsoInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("size_data_per_element", C.c_uint64), ("maxM", C.c_uint64),
                ("maxM0", C.c_uint64), ("M", C.c_uint64), ("ef_construction", C.c_uint64), ("dim", C.c_uint64),
                ("maxlevel", C.c_int32), ("threshold_level", C.c_int32), ("enterpoint", C.c_uint32),
                ("has_deleted", C.c_int32)]


class _HsoqInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("size_data_per_element", C.c_uint64), ("maxM", C.c_uint64),
                ("maxM0", C.c_uint64), ("M", C.c_uint64), ("ef_construction", C.c_uint64), ("dim", C.c_uint64),
                ("padded_dim", C.c_uint64), ("num_cluster", C.c_uint64), ("ex_bits", C.c_uint64),
                ("maxlevel", C.c_int32), ("threshold_level", C.c_int32), ("enterpoint", C.c_uint32),
                ("metric_type", C.c_int32)]


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False, oracle=True)
        L = C.CDLL(ORACLE_SO)
        L.hso_last_error.restype = C.c_char_p
        L.hso_load.restype = C.c_void_p
        L.hso_load.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.hso_load_hnsw.restype = C.c_void_p
        L.hso_load_hnsw.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.hso_free.argtypes = [C.c_void_p]
        L.hso_patch.restype = C.c_int
        L.hso_patch.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.hso_get_info.argtypes = [C.c_void_p, C.POINTER(_HsoInfo)]
        L.hso_node_level.restype = C.c_int
        L.hso_node_level.argtypes = [C.c_void_p, C.c_uint32]
        L.hso_node_label.restype = C.c_uint64
        L.hso_node_label.argtypes = [C.c_void_p, C.c_uint32]
        L.hso_node_vector.restype = C.POINTER(C.c_float)
        L.hso_node_vector.argtypes = [C.c_void_p, C.c_uint32]
        L.hso_node_neighbors.restype = C.c_int
        L.hso_node_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.POINTER(C.c_uint32))]
        L.hso_dist.restype = C.c_float
        L.hso_dist.argtypes = [_f32p, _f32p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.hso_search.restype = C.c_int
        L.hso_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                 _u32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hso_search_ties.restype = C.c_int
        L.hso_search_ties.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      _u32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hso_search_pool.restype = C.c_int
        L.hso_search_pool.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                      _u32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hso_bruteforce.restype = C.c_int
        L.hso_bruteforce.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, _f32p, C.c_size_t,
                                     C.c_size_t, C.c_int, _u32p, C.c_void_p]
        L.hso_recall.restype = C.c_double
        L.hso_recall.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _u32p, C.c_size_t, _u32p, C.c_size_t, C.c_int]
        L.hsoq_last_error.restype = C.c_char_p
        L.hsoq_load.restype = C.c_void_p
        L.hsoq_load.argtypes = [C.c_char_p, C.c_size_t]
        L.hsoq_free.argtypes = [C.c_void_p]
        L.hsoq_get_info.argtypes = [C.c_void_p, C.POINTER(_HsoqInfo)]
        L.hsoq_set_tconst.argtypes = [C.c_void_p, C.c_double]
        L.hsoq_node.restype = C.c_int
        L.hsoq_node.argtypes = [C.c_void_p, C.c_uint32, _u32p, _u64p, _f32p, _u32p, C.c_int]
        L.hsoq_rotate.argtypes = [C.c_void_p, _f32p, _f32p]
        L.hsoq_quantize_query.argtypes = [C.c_void_p, _f32p, _u64p, _f32p, C.c_void_p]
        L.hsoq_prep.argtypes = [C.c_void_p, _f32p, _f32p, _u64p, _f32p, _f32p]
        L.hsoq_est.restype = C.c_float
        L.hsoq_est.argtypes = [C.c_void_p, C.c_uint32, _u64p, _f32p, _f32p]
        L.hsoq_search.restype = C.c_int
        L.hsoq_search.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _u32p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
        _oracle_lib = L
    return _oracle_lib


class Oracle:
    """oracle/hs_oracle.c: the plain-C restatement of HierarchicalNSWSlim load + searchKnn."""

    def __init__(self, path: str, dim: int, metric: int = 0, hnsw: bool = False):
        """hnsw=True: `path` is the un-pruned hnswlib index of the `hnsw` strategy (hso_load_hnsw)."""
        self.L = oracle_lib()
        self.h = (self.L.hso_load_hnsw if hnsw else self.L.hso_load)(path.encode(), dim, metric)
        if not self.h:
            raise RuntimeError(self.L.hso_last_error().decode())
        self.dim, self.metric = dim, metric

    def close(self):
        if self.h:
            self.L.hso_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def patch(self, stream: bytes, rows: np.ndarray | None = None, inline: bool = False) -> None:
        """hso_patch: the restatement of patchFromStream (slim.h:2206-2388); rows[label] = vectors of new nodes."""
        ptr, n_rows = None, 0
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            ptr, n_rows = rows.ctypes.data_as(C.c_void_p), rows.shape[0]
        if self.L.hso_patch(self.h, stream, len(stream), 2 if inline else 0, ptr, n_rows) != 0:
            raise RuntimeError(self.L.hso_last_error().decode())

    def info(self) -> dict:
        s = _HsoInfo()
        self.L.hso_get_info(self.h, C.byref(s))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def node(self, i: int, level: int):
        p = C.POINTER(C.c_uint32)()
        cnt = self.L.hso_node_neighbors(self.h, i, level, C.byref(p))
        ids = np.ctypeslib.as_array(p, shape=(cnt,)).copy() if cnt else np.zeros(0, np.uint32)
        return self.L.hso_node_level(self.h, i), self.L.hso_node_label(self.h, i), ids

    def vector(self, i: int) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.hso_node_vector(self.h, i), shape=(self.dim,)).copy()

    def search(self, q, k: int, ef: int, order: int = ORDER_GPU, team: int = 8, threads: int = 0):
        """-> labels[nq,k], dists[nq,k] sorted by (dist, internal id); n_dist[nq], n_hops[nq]."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        lab = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        nd = np.zeros(nq, np.uint32)
        nh = np.zeros(nq, np.uint32)
        self.L.hso_search(self.h, q, nq, k, ef, order, team, threads, lab, dist.ctypes.data, nd.ctypes.data,
                          nh.ctypes.data)
        return lab, dist, nd, nh

    def search_pool(self, q, k: int, ef: int, team: int = 8, threads: int = 0):
        """hso_search_pool: the ENGINE's single-pool algorithm restated on the CPU -> labels, dists, n_dist, n_hops,
        n_ghosts (low 16 bits: ghosts expanded; high 16 bits: same-column tie displacements; oracle/hs_oracle.h)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        lab = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        nd, nh, ng = np.zeros(nq, np.uint32), np.zeros(nq, np.uint32), np.zeros(nq, np.uint32)
        if self.L.hso_search_pool(self.h, q, nq, k, ef, team, threads, lab, dist.ctypes.data, nd.ctypes.data,
                                  nh.ctypes.data, ng.ctypes.data) != 0:
            raise RuntimeError(self.L.hso_last_error().decode())
        return lab, dist, nd, nh, ng

    def search_ties(self, q, k: int, ef: int, order: int = ORDER_GPU, team: int = 8, threads: int = 0):
        """search() + n_ties[nq]: exact-tie events at the ef boundary per query (hso_search_ties)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        lab = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        nd, nh, nt = np.zeros(nq, np.uint32), np.zeros(nq, np.uint32), np.zeros(nq, np.uint32)
        self.L.hso_search_ties(self.h, q, nq, k, ef, order, team, threads, lab, dist.ctypes.data, nd.ctypes.data,
                               nh.ctypes.data, nt.ctypes.data)
        return lab, dist, nd, nh, nt


class OracleQ:
    """oracle/hs_oracle_slimq.c: the plain-C restatement of HierarchicalNSWSlimQ load + searchKnn.
    `base` are the raw rows the reference reranks with (setDataset, slimq.h:303-305)."""

    def __init__(self, path: str, base: np.ndarray, t_const: float | None = None):
        self.L = oracle_lib()
        self.base = np.ascontiguousarray(base, dtype=np.float32)
        self.dim = self.base.shape[1]
        self.h = self.L.hsoq_load(path.encode(), self.dim)
        if not self.h:
            raise RuntimeError(self.L.hsoq_last_error().decode())
        s = _HsoqInfo()
        self.L.hsoq_get_info(self.h, C.byref(s))
        self.info = {f: getattr(s, f) for f, _ in s._fields_}
        if t_const is not None:
            self.set_tconst(t_const)

    def close(self):
        if self.h:
            self.L.hsoq_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_tconst(self, t: float):
        self.L.hsoq_set_tconst(self.h, float(t))

    def node(self, i: int):
        cl = np.zeros(1, np.uint32)
        code = np.zeros(self.info["padded_dim"] // 64, np.uint64)
        fac = np.zeros(3, np.float32)
        nb = np.zeros(256, np.uint32)
        cnt = self.L.hsoq_node(self.h, i, cl, code, fac, nb, 256)
        return int(cl[0]), code, fac, nb[:cnt].copy()

    def rotate(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        out = np.zeros((q.shape[0], self.info["padded_dim"]), np.float32)
        for i in range(q.shape[0]):
            self.L.hsoq_rotate(self.h, q[i], out[i])
        return out

    def quantize(self, rotated):
        """-> planes[nq, pd/64*4], scal[nq,3], codes[nq,pd] (4-bit code per dimension)"""
        pd = self.info["padded_dim"]
        rotated = np.ascontiguousarray(rotated, dtype=np.float32).reshape(-1, pd)
        nq = rotated.shape[0]
        planes = np.zeros((nq, pd // 64 * 4), np.uint64)
        scal = np.zeros((nq, 3), np.float32)
        codes = np.zeros((nq, pd), np.uint16)
        for i in range(nq):
            self.L.hsoq_quantize_query(self.h, rotated[i], planes[i], scal[i], codes[i].ctypes.data)
        return planes, scal, codes

    def prep(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        pd, nc = self.info["padded_dim"], self.info["num_cluster"]
        nq = q.shape[0]
        rot = np.zeros((nq, pd), np.float32)
        planes = np.zeros((nq, pd // 64 * 4), np.uint64)
        scal = np.zeros((nq, 3), np.float32)
        q2c = np.zeros((nq, nc), np.float32)
        for i in range(nq):
            self.L.hsoq_prep(self.h, q[i], rot[i], planes[i], scal[i], q2c[i])
        return rot, planes, scal, q2c

    def est(self, ids, planes, scal, q2c):
        return np.array([self.L.hsoq_est(self.h, int(i), planes, scal, q2c) for i in ids], np.float32)

    def search(self, q, k: int, ef: int, order: int = ORDER_GPU, team: int = 8, threads: int = 0, inject=None):
        """-> labels[nq,k], dists[nq,k] (sorted by (dist,id)), n_est[nq], n_hops[nq], n_rerank[nq].
        inject = (planes, scal, q2c) replaces the oracle's own per-query preparation."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        lab = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        ne, nh, nr = (np.zeros(nq, np.uint32) for _ in range(3))
        ip = isc = iq = None
        if inject is not None:
            pl, sc, qc = (np.ascontiguousarray(a) for a in inject)
            assert pl.dtype == np.uint64 and sc.dtype == np.float32 and qc.dtype == np.float32
            ip, isc, iq = pl.ctypes.data, sc.ctypes.data, qc.ctypes.data
        rc = self.L.hsoq_search(self.h, self.base, q, nq, k, ef, order, team, threads, ip, isc, iq, lab,
                                dist.ctypes.data, ne.ctypes.data, nh.ctypes.data, nr.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.L.hsoq_last_error().decode())
        return lab, dist, ne, nh, nr


def oracle_dist(a, b, metric=0, order=ORDER_GPU, team=8) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(oracle_lib().hso_dist(a, b, a.shape[0], metric, order, team))


def oracle_bruteforce(base, q, k, metric=0, order=ORDER_REF, team=8, threads=0):
    base = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = min(k, base.shape[0])
    lab = np.zeros((q.shape[0], k), np.uint32)
    dist = np.zeros((q.shape[0], k), np.float32)
    oracle_lib().hso_bruteforce(base, base.shape[0], base.shape[1], metric, order, team, q, q.shape[0], k, threads,
                                lab, dist.ctypes.data)
    return lab, dist


def oracle_recall(base, q, knn, gt, K=None, metric=0) -> float:
    base = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    knn = np.ascontiguousarray(knn, dtype=np.uint32)
    gt = np.ascontiguousarray(gt, dtype=np.uint32)
    K = K or knn.shape[1]
    assert knn.shape[1] == K and gt.shape[1] >= K
    return float(oracle_lib().hso_recall(base, base.shape[1], q, q.shape[0], knn, K, gt, gt.shape[1], metric))


if __name__ == "__main__":
    import json
    import sys
    if len(sys.argv) == 3 and sys.argv[1] == "--build-slim":
        td = sys.argv[2]
        with open(os.path.join(td, "args.json")) as f:
            a = json.load(f)
        base = np.load(os.path.join(td, "base.npy"), mmap_mode="r")
        labels = np.load(os.path.join(td, "labels.npy")) if a["have_labels"] else None
        out = ref_slim_build(np.ascontiguousarray(base), a["path"], metric=a["metric"], M=a["M"],
                             ef_construction=a["ef_construction"], branching=a["branching"], threads=a["threads"],
                             labels=labels, hnsw_path=a["hnsw_path"], isolate=False, **a["prune"])
        print(json.dumps(out))
    if len(sys.argv) == 3 and sys.argv[1] == "--make-patches":
        td = sys.argv[2]
        with open(os.path.join(td, "args.json")) as f:
            a = json.load(f)
        ref_slim_make_patches(np.ascontiguousarray(np.load(os.path.join(td, "base.npy"))), a["n0"], a["rounds"],
                              a["partial_path"], a["patch_prefix"], final_path=a["final_path"],
                              inline_last=a["inline_last"], metric=a["metric"], M=a["M"],
                              ef_construction=a["ef_construction"], branching=a["branching"], threads=a["threads"],
                              isolate=False, **a["prune"])
    if len(sys.argv) == 3 and sys.argv[1] == "--build-slimq":
        td = sys.argv[2]
        with open(os.path.join(td, "args.json")) as f:
            a = json.load(f)
        out = ref_slimq_build(np.load(os.path.join(td, "base.npy")), np.load(os.path.join(td, "cent.npy")),
                              np.load(os.path.join(td, "cid.npy")), a["path"], M=a["M"],
                              ef_construction=a["ef_construction"], threads=a["threads"], isolate=False,
                              **a["prune"])
        print(json.dumps(out))
    if len(sys.argv) == 3 and sys.argv[1] in ("--build-hnsw", "--build-slimzero"):
        td = sys.argv[2]
        with open(os.path.join(td, "args.json")) as f:
            a = json.load(f)
        base = np.ascontiguousarray(np.load(os.path.join(td, "base.npy")))
        if sys.argv[1] == "--build-hnsw":
            ref_hnsw_build(base, a["path"], metric=a["metric"], M=a["M"], ef_construction=a["ef_construction"],
                           branching=a["branching"], threads=a["threads"], isolate=False)
        else:
            ref_slimzero_build(base, a["path"], metric=a["metric"], M=a["M"], ef_construction=a["ef_construction"],
                               branching=a["branching"], threads=a["threads"], min_indegree0=a["min_indegree0"],
                               min_indegree=a["min_indegree"], isolate=False, **a["prune"])
