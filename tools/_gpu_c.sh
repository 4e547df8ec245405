mkdir -p gpurun_out/r2b
(time python -m pytest tests/test_gpu_traverse.py tests/test_gpu_slimq.py -m gpu -q -k "ef_129 or compact or slimq" 2>&1 | tail -15) > gpurun_out/r2b/gputests_c.log 2>&1
P="python tools/perf_probe.py --gpu-build 1 --overlap 1 --iters 20 --check 1"
$P --workload sift1m --efs 50,100,200 > gpurun_out/r2b/ring_default.txt 2>&1
HS_LIB_PATH=hnsw_slim_b200/_build/alt_ring8/libhnswslim_b200.so $P --workload sift1m --efs 50,100,200 > gpurun_out/r2b/ring_8.txt 2>&1
HS_LIB_PATH=hnsw_slim_b200/_build/alt_ring12/libhnswslim_b200.so $P --workload sift1m --efs 50,100,200 > gpurun_out/r2b/ring_12.txt 2>&1
$P --workload msturing1m-slimq --efs 50,100,200 > gpurun_out/r2b/slimq_est32.txt 2>&1
HS_LIB_PATH=hnsw_slim_b200/_build/alt_slimq_old/libhnswslim_b200.so $P --workload msturing1m-slimq --efs 50,100,200 > gpurun_out/r2b/slimq_old.txt 2>&1
HS_LIB_PATH=hnsw_slim_b200/_build/alt_slimq_top/libhnswslim_b200.so $P --workload msturing1m-slimq --efs 50,100,200 > gpurun_out/r2b/slimq_top.txt 2>&1
