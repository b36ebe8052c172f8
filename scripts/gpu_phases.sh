mkdir -p gpurun_out
{
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config2 --chain-steps 32
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config2 --chain-steps 32 --chains 296
RECOM_B200_LIB=$PWD/variants/lib_timers.so python scripts/phases.py --workload config2 --chain-steps 32 --chains 1184
} > gpurun_out/phases.log 2>&1
