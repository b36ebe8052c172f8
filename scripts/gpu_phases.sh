mkdir -p gpurun_out
{
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config2
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config3
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config3 --wilson --chain-steps 4
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config3 --cap-nodes 9000 --cap-edges 26976 --chain-steps 4
} > gpurun_out/phases.log 2>&1
python bench.py --workload config3 --wilson --chain-steps 4 --steps 5 --warmup 2 --cpu-seconds 8 > gpurun_out/bench_c3w.json 2> gpurun_out/bench_c3w.err
python bench.py --workload config1 --wilson --chain-steps 8 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c1w.json 2>> gpurun_out/bench_c3w.err
