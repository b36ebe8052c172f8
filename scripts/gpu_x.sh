mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_26.log
rm -f gpurun_out/ab.log gpurun_out/ab.err
bash scripts/gpu_ab.sh base w
