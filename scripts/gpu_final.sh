# Round-end evidence run on ONE box: full GPU tests, smoke, every bench line, ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; echo "tests rc=$?" > gpurun_out/final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final.log
B="timeout 600 python bench.py"
$B > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
$B --impl reference > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1.err
$B --workload config3 --no-cpu-baseline > gpurun_out/bench_r1_c3.json 2>> gpurun_out/bench_r1.err
$B --workload config5 --no-cpu-baseline > gpurun_out/bench_r1_c5.json 2>> gpurun_out/bench_r1.err
$B --workload config1 --no-cpu-baseline > gpurun_out/bench_r1_c1.json 2>> gpurun_out/bench_r1.err
$B --workload config3 --wilson --chain-steps 128 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_r1_c3w.json 2>> gpurun_out/bench_r1.err
$B --workload config1 --wilson --chain-steps 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c1w.json 2>> gpurun_out/bench_r1.err
$B --workload config4 --wilson --chain-steps 16 --steps 2 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_r1_c4.json 2>> gpurun_out/bench_r1.err
CMD="python bench.py --steps 3 --warmup 3 --chain-steps 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches_r1.log 2>&1
echo done >> gpurun_out/final.log
