"""The C++ host layer (hnsw_slim_b200/host/: SolveStrategy-shaped classes + the reference's command
line, main.cc:10-139) above the C ABI.  CPU: flag surface, index naming, loud failure without a
GPU.  GPU: the three strategies end to end on .fvecs files, against the Python binding."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU
from hnsw_slim_b200 import build as hs_build
from hnsw_slim_b200 import capi, vecs_io
from hnsw_slim_b200.synth import make_dataset


def run_cli(*args, cwd=None):
    r = subprocess.run([hs_build.HOST_CLI, *args], capture_output=True, text=True, cwd=cwd, timeout=600)
    return r.returncode, r.stdout, r.stderr


def make_files(tmp_path, name="toy", n=20000, nq=200, dim=96):
    base, q = make_dataset(n, nq, dim, rank=10, seed=9)
    d = tmp_path / "data" / name
    d.mkdir(parents=True)
    vecs_io.write_vecs(str(d / f"{name}_base.fvecs"), base)
    vecs_io.write_vecs(str(d / f"{name}_query.fvecs"), q)
    return base, q, str(tmp_path / "data"), str(tmp_path / "index")


def test_unknown_strategy_and_flags(tmp_path):
    rc, out, _ = run_cli("--solve_strategy=annoy", "--data_dir", str(tmp_path))
    assert rc == 1 and "Unknown strategy: annoy" in out                      # main.cc:134-138
    rc, out, _ = run_cli("--no_such_flag=1")
    assert rc == 1 and "unknown command line flag" in out
    # index file name and derived pruning parameters, main.cc:58-100
    rc, out, _ = run_cli("--solve_strategy=nope", "--dataset=sift", "--m=16", "--ef_construction=200",
                         "--branching_factor=4", "--top_M0=32", "--Mm_ratio=25", "--level_ratio=50")
    assert "Index path: ../statistics/index/sift/nope_200_16_4_0_0.020000_0.020000_32_8_16_4.graph" in out
    assert "top_m0: 32, top_m: 16, low_m0: 8, low_m: 4" in out


def test_missing_dataset_exits_like_the_reference(tmp_path):
    rc, out, _ = run_cli("--solve_strategy=hnsw_slim", "--dataset=nothing", "--data_dir", str(tmp_path))
    assert rc != 0 and "open file error" in out                             # util.h:57-60: exit(-1)


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU behaviour")
def test_no_gpu_fails_loudly(tmp_path):
    base, q, data_dir, index_dir = make_files(tmp_path, n=2000, nq=10, dim=32)
    rc, out, err = run_cli("--solve_strategy=hnsw_slim", "--dataset=toy", "--data_dir", data_dir, "--index_dir",
                           index_dir, "--m=8", "--ef_construction=40", "--k=5")
    assert rc != 0
    assert "no CUDA device" in err and "no CPU fallback" in err
    # the index was still built and saved on the host, in the reference's format
    graphs = list((tmp_path / "index" / "toy").glob("hnsw_slim_40_8_4_*.graph"))
    assert len(graphs) == 1 and capi.HostGraph(str(graphs[0]), 32).info()["n"] == 2000


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    base, q, data_dir, index_dir = make_files(tmp_path)
    common = ["--dataset=toy", "--data_dir", data_dir, "--index_dir", index_dir, "--m=16", "--ef_construction=100",
              "--k=10", "--ef_search=100"]
    # 1. ground truth (brute_force_strategy.h:15-45): k=100 rows, farthest first
    rc, out, err = run_cli("--solve_strategy=bruteforce", *common)
    assert rc == 0, err
    gt = vecs_io.read_ivecs(os.path.join(data_dir, "toy", "toy_groundtruth.ivecs"))
    assert gt.shape == (200, 100)
    want, _ = capi.bruteforce_knn(base, q, 100)
    assert np.array_equal(gt, want[:, ::-1])
    assert float(re.search(r"Recall: ([0-9.]+)", out).group(1)) == 1.0        # its own top-10 against itself
    # 2. hnsw_slim: builds + saves on the first run, loads on the second; same recall both times
    recalls = []
    for _ in range(2):
        rc, out, err = run_cli("--solve_strategy=hnsw_slim", *common)
        assert rc == 0, err
        recalls.append(float(re.search(r"Recall: ([0-9.]+)", out).group(1)))
    assert recalls[0] == recalls[1] and recalls[0] >= 0.95
    graph = [p for p in os.listdir(os.path.join(index_dir, "toy")) if p.startswith("hnsw_slim_100_16_4_")]
    assert len(graph) == 1
    ix = capi.Index(os.path.join(index_dir, "toy", graph[0]), 96)
    ix.set_ef(100)
    lab, _ = ix.search(q, 10)
    py_recall = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(lab, want[:, :10])])
    assert abs(py_recall - recalls[0]) < 1e-6
    # 3. hnsw-slimq (README spelling)
    rc, out, err = run_cli("--solve_strategy=hnsw-slimq", *common)
    assert rc == 0, err
    assert float(re.search(r"Recall: ([0-9.]+)", out).group(1)) >= 0.93
    assert any(p.startswith("hnsw_slimq_100_16_4_") for p in os.listdir(os.path.join(index_dir, "toy")))


def test_hnsw_builder_writes_the_reference_format(tmp_path):
    """hs_build_hnsw_graph -> HierarchicalNSW::saveIndex's format: the engine's loader, the oracle and
    (where it is available) the reference's own loadIndex read the file; searching it with the
    reference gives a working index (recall)."""
    from oracle import refharness as rh
    base, q = make_dataset(5000, 100, 24, rank=8, seed=5)
    g = str(tmp_path / "h.graph")
    capi.build_hnsw_graph(base, g, M=12, ef_construction=80)
    info = capi.HostGraph(g, 24, kind=capi.HS_KIND_HNSW).info()
    assert (info["n"], info["maxM"], info["maxM0"], info["M"], info["threshold_level"]) == (5000, 12, 24, 12, 0)
    orc = rh.Oracle(g, 24, 0, hnsw=True)
    ol, _, _, _ = orc.search(q, 10, 80, order=rh.ORDER_REF)
    gt, _ = rh.oracle_bruteforce(base, q, 10)
    rec = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ol, gt)])
    assert rec >= 0.95, rec
    if rh.ref_hnsw_path():
        rl, _, _ = rh.RefHnsw(g, 24, 5000).search(q, 10, 80)
        assert np.mean([set(a) == set(b) for a, b in zip(ol, rl)]) >= 0.99


@pytest.mark.gpu
def test_cli_hnsw_and_slimzero(tmp_path):
    from oracle import refharness as rh
    base, q, data_dir, index_dir = make_files(tmp_path, n=10000, nq=100, dim=32)
    common = ["--dataset=toy", "--data_dir", data_dir, "--index_dir", index_dir, "--m=16", "--ef_construction=100",
              "--k=10", "--ef_search=100"]
    rc, out, err = run_cli("--solve_strategy=bruteforce", *common)
    assert rc == 0, err
    recalls = []
    for _ in range(2):                                   # builds + saves, then loads (hnsw_strategy.h:21-45)
        rc, out, err = run_cli("--solve_strategy=hnsw", *common)
        assert rc == 0, err
        recalls.append(float(re.search(r"Recall: ([0-9.]+)", out).group(1)))
    assert recalls[0] == recalls[1] and recalls[0] >= 0.97
    assert any(p.startswith("hnsw_100_16_4_") for p in os.listdir(os.path.join(index_dir, "toy")))
    # hnsw_slimzero: load-only — without the file a loud error, with a reference-built file a search
    rc, out, err = run_cli("--solve_strategy=hnsw-slimzero", *common)
    assert rc != 0 and "build it with the reference" in err
    if rh.ref_hnsw_path():
        path = re.search(r"Index path: (\S+)", out).group(1)
        rh.ref_slimzero_build(base, path, M=16, ef_construction=100)
        rc, out, err = run_cli("--solve_strategy=hnsw-slimzero", *common)
        assert rc == 0, err
        assert float(re.search(r"Recall: ([0-9.]+)", out).group(1)) >= 0.9


@pytest.mark.gpu
def test_cli_sharded_mode_one_process(tmp_path):
    """hs_main --shards S --gpus N: the C++ host path of the sharded corpus (no Python, no torch.distributed): the
    sub-graphs are built on the GPU(s), one host thread drives one hs_shardgroup per GPU over peer memory, batches
    of --batch queries stream through.  On a one-GPU box N = 1 (all shards local); with more devices N = 2."""
    import torch
    base, q, data_dir, index_dir = make_files(tmp_path, n=24000, nq=300, dim=96)
    common = ["--dataset=toy", "--data_dir", data_dir, "--index_dir", index_dir, "--m=16", "--ef_construction=100",
              "--k=10"]
    rc, out, err = run_cli("--solve_strategy=bruteforce", *common)
    assert rc == 0, err
    gpus = 2 if torch.cuda.device_count() >= 2 else 1
    rc, out, err = run_cli("--solve_strategy=hnsw_slim", "--shards=4", f"--gpus={gpus}", "--batch=64", "--ef_search=40",
                           *common)
    assert rc == 0, err
    assert f"for 4 shards on {gpus} GPUs" in out
    assert float(re.search(r"Recall: ([0-9.]+)", out).group(1)) >= 0.97
    rc, out, err = run_cli("--solve_strategy=hnsw_slim", "--shards=3", "--gpus=2", *common)
    assert rc != 0 and "multiple of --gpus" in err


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU behaviour")
def test_serve_strategy_without_a_gpu_fails_loudly(tmp_path):
    base, q, data_dir, index_dir = make_files(tmp_path, n=1200, nq=10, dim=16)
    graph = os.path.join(os.path.dirname(__file__), "golden", "patch_l2_1k.graph")
    rc, out, err = run_cli("--solve_strategy=hnsw_slim_serve", "--dataset=toy", "--data_dir", data_dir, "--index_path", graph,
                           "--k=5", "--threads=4")
    assert rc != 0 and "no CUDA device" in err and "no CPU fallback" in err


@pytest.mark.gpu
def test_cli_serve_with_patches(tmp_path):
    """hs_main --solve_strategy=hnsw_slim_serve: the reference client's flow (partial index opened with room for
    the whole base set, the server's patch streams applied, hnsw_slim_client_update_patch.cc:113,147-179) with the
    query set going through the serving front one query per call from 16 threads."""
    from conftest import HAVE_REF
    from oracle import refharness as rh
    if not HAVE_REF:
        pytest.skip("oracle/_ref (compiled reference) not available on this host")
    base, q, data_dir, index_dir = make_files(tmp_path, n=12000, nq=300, dim=64)
    part = str(tmp_path / "part.graph")
    names = rh.ref_slim_make_patches(base, 9000, 2, part, str(tmp_path / "p"), inline_last=True, M=16,
                                     ef_construction=100, threads=1)
    common = ["--dataset=toy", "--data_dir", data_dir, "--index_dir", index_dir, "--k=10", "--ef_search=64"]
    rc, out, err = run_cli("--solve_strategy=bruteforce", *common)
    assert rc == 0, err
    rc, out, err = run_cli("--solve_strategy=hnsw_slim_serve", *common, "--index_path", part, "--patches",
                           ",".join(names), "--patch_inline_last", "--threads=16", "--max_batch=64")
    assert rc == 0, err
    assert "9000 -> 10500 elements" in out and "10500 -> 12000 elements" in out
    m = re.search(r"served (\d+) queries from 16 threads in (\d+) batches \(largest (\d+)\)", out)
    assert m and int(m.group(1)) == 300 and int(m.group(2)) < 300 and 1 < int(m.group(3)) <= 64
    got = float(re.search(r"Recall: ([0-9.]+)", out).group(1))
    ix = capi.Index.load_reserve(part, 64, 12000)
    ix.patch(open(names[0], "rb").read(), rows=base)
    ix.patch(open(names[1], "rb").read(), inline=True)
    ix.set_ef(64)
    lab, _ = ix.search(q, 10)
    gt, _ = capi.bruteforce_knn(base, q, 10)
    want = float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(lab, gt)]))
    assert abs(got - want) < 1e-6 and got > 0.9


def test_host_util_reads_and_writes_vecs(tmp_path):
    """host/util.h (ReadData / WriteData of .fvecs / .ivecs: one image per file, row headers checked) round-trips
    files written by the Python side byte for byte, across its 8 MB slab boundary, and rejects a bad row header."""
    host = os.path.join(os.path.dirname(hs_build.HERE), "hnsw_slim_b200", "host")
    src = tmp_path / "t.cc"
    src.write_text('#include "util.h"\n'
                   'int main(int argc, char **argv) {\n'
                   '  std::vector<float> rows; uint32_t num = 0, dim = 0;\n'
                   '  ReadData(argv[1], rows, num, dim);\n'
                   '  WriteData(argv[2], rows, num, dim);\n'
                   '  return 0;\n}\n')
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", host, "-o", exe, str(src)], check=True)
    rng = np.random.default_rng(0)
    for n, d in ((1000, 33), (6000, 960), (1, 1)):
        a = rng.standard_normal((n, d)).astype(np.float32)
        fin, fout = str(tmp_path / f"in_{n}.fvecs"), str(tmp_path / f"out_{n}.fvecs")
        vecs_io.write_vecs(fin, a)
        r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
        assert r.returncode == 0 and f"num: {n}" in r.stdout and f"dim: {d}" in r.stdout
        assert open(fin, "rb").read() == open(fout, "rb").read()
    bad = np.fromfile(str(tmp_path / "in_1000.fvecs"), np.int32)
    bad[5 * 34] = 32                                          # the header of row 5
    bad.tofile(str(tmp_path / "bad.fvecs"))
    r = subprocess.run([exe, str(tmp_path / "bad.fvecs"), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0 and "open file error" in r.stdout and "row 5" in r.stdout
    r = subprocess.run([exe, str(tmp_path / "missing.fvecs"), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0 and "open file error" in r.stdout
