# plain vs shared-retry build after the register work (config 2 / 3 / 5)
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -2
run() { echo "$1 $2 S=$3: $(env $1 timeout 120 python scripts/dev/quick.py $2 --chains 1024 --S $3 --reps $4 2>&1 | tail -1 | sed 's/.*best/best/')"; }
run "RECOM_B200_SHARE=0" config2 256 4
run "RECOM_B200_SHARE=3 RECOM_B200_HELP_AFTER=100000000" config2 256 4
run "RECOM_B200_SHARE=1" config2 256 4
run "RECOM_B200_SHARE=0" config3 512 8
run "RECOM_B200_SHARE=1" config3 512 8
run "RECOM_B200_SHARE=1" config3 32 8
run "RECOM_B200_SHARE=1" config5 512 6
