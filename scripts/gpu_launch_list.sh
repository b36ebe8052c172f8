# ncu launch list of the bench command (after it has exited 0 without ncu): bash scripts/gpu_launch_list.sh
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --chain-steps 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches_r1.log 2>&1
