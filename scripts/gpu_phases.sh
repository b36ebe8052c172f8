mkdir -p gpurun_out
{
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config3 --wilson --chain-steps 16
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config1 --wilson --chain-steps 32
} > gpurun_out/phases.log 2>&1
