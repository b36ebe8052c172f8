# recruiting + keep: parity, then mean/best throughput with share on/off
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -3
for SH in 1 0; do
  for S in 32 512; do
    echo "share=$SH S=$S: $(RECOM_B200_SHARE=$SH timeout 120 python scripts/dev/quick.py config3 --chains 1024 --S $S --reps 6 2>&1 | tail -1)"
  done
  echo "share=$SH config2 S=512: $(RECOM_B200_SHARE=$SH timeout 120 python scripts/dev/quick.py config2 --chains 1024 --S 512 --reps 4 2>&1 | tail -1)"
done
echo "share=1 config5 S=512: $(RECOM_B200_SHARE=1 timeout 120 python scripts/dev/quick.py config5 --chains 1024 --S 512 --reps 6 2>&1 | tail -1)"
echo "share=1 wilson S=128: $(RECOM_B200_SHARE=1 timeout 200 python scripts/dev/quick.py config3 --wilson --chains 1024 --S 128 --reps 3 2>&1 | tail -1)"
echo "share=1 592 chains S=512: $(RECOM_B200_SHARE=1 timeout 200 python scripts/dev/quick.py config3 --chains 592 --S 512 --reps 4 2>&1 | tail -1)"
