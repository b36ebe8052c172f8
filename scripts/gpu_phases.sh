mkdir -p gpurun_out
{
echo "MST build: slots 11-14 = boruvka weights+init | A+B | C+C' | D"
RECOM_B200_LIB=$PWD/variants/lib_timers_b.so python scripts/phases.py --workload config2 --chain-steps 64
RECOM_B200_LIB=$PWD/variants/lib_timers_b.so python scripts/phases.py --workload config3 --chain-steps 64 --chains 1024
} > gpurun_out/phases_b.log 2>&1
