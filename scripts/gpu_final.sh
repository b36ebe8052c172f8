# Round-end check on ONE box: full GPU tests, smoke, the default bench line (+ config3).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; echo "tests rc=$?" > gpurun_out/final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final.log
B="timeout 600 python bench.py"
$B > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
$B --workload config3 --no-cpu-baseline > gpurun_out/bench_r1_c3.json 2>> gpurun_out/bench_r1.err
echo done >> gpurun_out/final.log
