set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python bench.py --steps 5 --warmup 3 --workload config3 --no-cpu-baseline > gpurun_out/bench_a_c3.json 2>> gpurun_out/bench_a.err
python bench.py --steps 5 --warmup 3 --threads-per-cta 512 --no-cpu-baseline > gpurun_out/bench_a_t512.json 2>> gpurun_out/bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_a_ref.json 2>> gpurun_out/bench_a.err
CMD="python bench.py --steps 2 --warmup 1 --chain-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 3 -c 1 -o gpurun_out/prof_a $CMD > gpurun_out/ncu2.log 2>&1
