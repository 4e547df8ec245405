mkdir -p gpurun_out/r2b
(python -m pytest tests/test_gpu_traverse.py -m gpu -q -k "shared_memory_pool or threshold_level" 2>&1 | tail -5) > gpurun_out/r2b/gputests_g.log 2>&1
P="python tools/perf_probe.py --gpu-build 1 --overlap 1 --iters 10 --check 1"
$P --workload gist200k --efs 300,400,512 --tflags 9,73 > gpurun_out/r2b/bigpool_default.txt 2>&1
HS_LIB_PATH=hnsw_slim_b200/_build/alt_huge5/libhnswslim_b200.so $P --workload gist200k --efs 300,400,512 --tflags 9 > gpurun_out/r2b/bigpool_huge5.txt 2>&1
B="python bench.py --steps 5 --warmup 3 --no-sharded --no-cpu-baseline --no-recall --sustained 0"
$B > gpurun_out/r2b/ncu_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b/launches_bench.csv $B > gpurun_out/r2b/ncu_bench.log 2>&1
Q="python tools/perf_probe.py --workload sift1m --gpu-build 1 --overlap 0 --iters 3 --efs 200"
$Q > gpurun_out/r2b/ncu_plain_probe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 3 -c 2 -o gpurun_out/r2b/prof_ef200 $Q > gpurun_out/r2b/ncu_probe.log 2>&1
