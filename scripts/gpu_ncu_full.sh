# one ncu --set full capture of the chain-step kernel (after the command has exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --chain-steps 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 4 -c 1 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full_r1.log 2>&1
