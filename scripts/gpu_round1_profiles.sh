# Round-1 evidence run: bench (both arms), ncu launch list, one ncu --set full capture.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1.err
python bench.py --workload config3 --no-cpu-baseline > gpurun_out/bench_r1_c3.json 2>> gpurun_out/bench_r1.err
python bench.py --workload config5 --no-cpu-baseline > gpurun_out/bench_r1_c5.json 2>> gpurun_out/bench_r1.err
python bench.py --workload config1 --no-cpu-baseline > gpurun_out/bench_r1_c1.json 2>> gpurun_out/bench_r1.err
CMD="python bench.py --steps 3 --warmup 3 --chain-steps 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches_r1.log 2>&1
$CMD > gpurun_out/plain_r1b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 4 -c 1 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full_r1.log 2>&1
