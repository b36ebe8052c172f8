# round 2: lane-walker Wilson path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wilson or reversible" > gpurun_out/w1_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/w1.log
for args in "config3 --wilson" "config3 --wilson --walkers 2" "config3 --wilson --walkers 8" "config3 --wilson --walkers 1" \
            "config3 --wilson --walkers 4 --threads 128" "config3 --wilson --walkers 3" "config3 --wilson --walkers 6" \
            "config1 --wilson" "config2 --wilson" \
            "config4 --S 8 --reps 2" "config4 --S 8 --reps 2 --walkers 2" "config4 --S 8 --reps 2 --walkers 8"; do
  timeout 300 python scripts/dev/quick.py $args >> gpurun_out/w1.log 2>> gpurun_out/w1.err
done
