# shared-retry build after the constant-memory Params / help bitmap change: parity, then share on/off
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "shared or free or wilson" 2>&1 | tail -3
for SH in 0 1; do
  for S in 64 512; do
    echo "share=$SH S=$S: $(RECOM_B200_SHARE=$SH timeout 120 python scripts/dev/quick.py config3 --chains 1024 --S $S --reps 4 2>&1 | tail -1)"
  done
  echo "share=$SH config2 S=256: $(RECOM_B200_SHARE=$SH timeout 120 python scripts/dev/quick.py config2 --chains 1024 --S 256 --reps 3 2>&1 | tail -1)"
done
echo "share=1 wilson S=64: $(RECOM_B200_SHARE=1 timeout 200 python scripts/dev/quick.py config3 --wilson --chains 1024 --S 64 --reps 2 2>&1 | tail -1)"
echo "share=0 wilson S=64: $(RECOM_B200_SHARE=0 timeout 200 python scripts/dev/quick.py config3 --wilson --chains 1024 --S 64 --reps 2 2>&1 | tail -1)"
