# ncu --set full capture of the Wilson chain-step kernel on a small launch (1 CTA per SM)
mkdir -p gpurun_out
CMD="python scripts/dev/quick.py config3 --wilson --walkers ${1:-1} --chains 148 --S 2 --reps 1"
$CMD > gpurun_out/plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 1 -c 1 -f -o gpurun_out/prof_w $CMD > gpurun_out/ncu_w.log 2>&1
tail -3 gpurun_out/ncu_w.log
