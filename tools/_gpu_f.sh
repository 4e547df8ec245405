mkdir -p gpurun_out/r2b
(time python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2b/gputests_f.log 2>&1
(time python bench.py --steps 20 --warmup 3 > gpurun_out/r2b/bench_default.json 2> gpurun_out/r2b/bench_default.err) 2>> gpurun_out/r2b/bench_default.err
(time python bench.py --workload bruteforce1m --steps 5 --warmup 3 > gpurun_out/r2b/bench_bruteforce1m.json 2> gpurun_out/r2b/bench_bruteforce1m.err) 2>> gpurun_out/r2b/bench_bruteforce1m.err
(time python bench.py --workload msturing1m-slimq --steps 20 --warmup 3 > gpurun_out/r2b/bench_msturing1m_slimq.json 2> gpurun_out/r2b/bench_msturing1m_slimq.err) 2>> gpurun_out/r2b/bench_msturing1m_slimq.err
(time python bench.py --workload cohere1m --steps 20 --warmup 3 > gpurun_out/r2b/bench_cohere1m.json 2> gpurun_out/r2b/bench_cohere1m.err) 2>> gpurun_out/r2b/bench_cohere1m.err
