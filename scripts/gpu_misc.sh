# Wilson phase split (RC_TIMERS build)
RECOM_B200_SHARE=0 RECOM_B200_LIB=variants/lib_timers.so timeout 200 python scripts/phases.py --workload config3 --wilson --chain-steps 64 2>&1 | tail -18
