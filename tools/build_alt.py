"""Developer helper: an alternative build of the library that differs from the main one in ONE translation unit
compiled with extra -D flags (kernel experiments), linked against the main build's other objects.
    python tools/build_alt.py <tag> <source.cu> <DEF[=V]>...   ->  hnsw_slim_b200/_build/alt_<tag>/libhnswslim_b200.so
Selected at run time with HS_LIB_PATH (the Python binding)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnsw_slim_b200 import build as b  # noqa: E402

tag, srcs, defs = sys.argv[1], [a for a in sys.argv[2:] if a.endswith((".cu", ".cpp"))], [a for a in sys.argv[2:] if not a.endswith((".cu", ".cpp"))]
b.build()
out = os.path.join(b.OUT_DIR, f"alt_{tag}")
os.makedirs(out, exist_ok=True)
objs = []
procs = []
for src in b.SOURCES:
    stem = os.path.splitext(src)[0]
    if src in srcs:
        obj = os.path.join(out, stem + ".o")
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-c", os.path.join(b.CSRC, src), "-o", obj]))
    else:
        obj = os.path.join(b.OUT_DIR, stem + ".o")
    objs.append(obj)
for p in procs:
    if p.wait() != 0:
        raise SystemExit("nvcc failed")
lib = os.path.join(out, "libhnswslim_b200.so")
subprocess.run([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"], check=True)
for o in objs:
    if o.startswith(out):
        os.remove(o)
print(lib)
