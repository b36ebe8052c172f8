timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_ensemble_ks.py -q -x -m gpu -k "wilson or reversible or spill or shared" 2>&1 | tail -3
run() { echo "$1 | $2: $(env $1 timeout 300 python scripts/dev/quick.py config3 --wilson --chains 1024 --S 128 --reps 4 $2 2>&1 | tail -1 | cut -c1-300)"; }
run "RECOM_B200_SHARE=0 RECOM_B200_NBR32=1" ""
run "RECOM_B200_SHARE=0" ""
run "RECOM_B200_SHARE=0" "--threads 192"
run "RECOM_B200_SHARE=1" ""
