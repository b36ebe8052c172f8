mkdir -p gpurun_out
S=$(date +%s)
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
echo "default bench wall $(( $(date +%s) - S )) s" > gpurun_out/bench_wall.log
S=$(date +%s)
python bench.py --impl reference > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1.err
echo "reference arm wall $(( $(date +%s) - S )) s" >> gpurun_out/bench_wall.log
python bench.py --workload config3 --no-cpu-baseline > gpurun_out/bench_r1_c3.json 2>> gpurun_out/bench_r1.err
python bench.py --workload config5 --no-cpu-baseline > gpurun_out/bench_r1_c5.json 2>> gpurun_out/bench_r1.err
python bench.py --workload config1 --no-cpu-baseline > gpurun_out/bench_r1_c1.json 2>> gpurun_out/bench_r1.err
