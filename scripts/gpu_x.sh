mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "wilson or config4 or spill or ensemble or edge_cases" 2>&1 | tail -8 > gpurun_out/gpu_tests_21.log
python bench.py --workload config3 --wilson --chain-steps 16 --steps 10 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3w3.json 2> gpurun_out/bench_c3w3.err
timeout 500 python bench.py --workload config4 --chains 1024 --chain-steps 4 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4c.json 2>> gpurun_out/bench_c3w3.err
python bench.py --workload config1 --wilson --no-cpu-baseline --chain-steps 32 --steps 20 > gpurun_out/bench_c1w2.json 2>> gpurun_out/bench_c3w3.err
