mkdir -p gpurun_out
B="timeout 100 python bench.py"
$B --workload config5 --no-cpu-baseline --steps 20 > gpurun_out/bench_r1_c5.json 2>> gpurun_out/bench_r1.err
$B --workload config3 --wilson --chain-steps 128 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c3w.json 2>> gpurun_out/bench_r1.err
$B --workload config1 --no-cpu-baseline --steps 20 > gpurun_out/bench_r1_c1.json 2>> gpurun_out/bench_r1.err
$B --workload config1 --wilson --chain-steps 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c1w.json 2>> gpurun_out/bench_r1.err
