"""Builds hnsw_slim_b200/_build/libhnswslim_b200.so (CUDA kernels + C ABI) in-tree with nvcc
for sm_100a.  nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhnswslim_b200.so")

SOURCES = ["hs_api.cu", "graph_loader.cpp", "graph_build.cpp", "traverse_fp32.cu", "traverse_fp32_g.cu", "traverse_fp32_c.cu", "traverse_slimq.cu", "bruteforce.cu", "bruteforce_tc.cu", "shard_group.cu", "graph_gpu.cu", "patch.cu", "service.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
              "-ccbin", "/usr/bin/g++"]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    host = os.path.join(HERE, "host")
    deps = ([os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(host, f) for f in os.listdir(host)]
            + [os.path.join(HERE, "..", "include", "hnswslim_b200.h")])
    if not os.path.exists(os.path.join(OUT_DIR, "hs_main")):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(tag: str, defs: list[str]) -> str:
    """Developer helper: an extra copy of the library compiled with -D flags (tuning sweeps);
    selected at run time with HS_LIB_PATH."""
    out = os.path.join(OUT_DIR, f"variant_{tag}")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libhnswslim_b200.so")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(out, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append(subprocess.Popen(cmd))
        objs.append(obj)
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.run([_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                            "-ccbin", "/usr/bin/g++"], check=True)
    return lib


def build(force: bool = False, verbose: bool = False, ptxas_v: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or ptxas_v:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"]
    subprocess.run(cmd, check=True)
    build_host_cli()
    return LIB


HOST_CLI = os.path.join(OUT_DIR, "hs_main")


def build_host_cli() -> str:
    """The C++ host layer above the C ABI: hnsw_slim_b200/host/ (SolveStrategy-shaped classes and the
    reference's command line) -> _build/hs_main, linked against the shared library next to it."""
    host = os.path.join(HERE, "host")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-o", HOST_CLI, os.path.join(host, "main.cc"),
           "-L" + OUT_DIR, "-lhnswslim_b200", "-pthread", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return HOST_CLI


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_v="--ptxas" in sys.argv))
