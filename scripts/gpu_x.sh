mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_16.log
rm -f gpurun_out/ab.log gpurun_out/ab.err
bash scripts/gpu_ab.sh o r
{
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config2 --chain-steps 32
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config3 --chain-steps 32
} > gpurun_out/phases.log 2>&1
