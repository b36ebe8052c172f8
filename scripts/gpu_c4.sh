timeout 500 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "spill or l2 or config4 or wilson" 2>&1 | tail -3
run() { echo "$1 | $2: $(env $1 timeout 300 python scripts/dev/quick.py config4 --chains 1024 --S 8 --reps 3 $2 2>&1 | tail -1 | cut -c1-260)"; }
run "RECOM_B200_SHARE=0" ""
run "" ""
