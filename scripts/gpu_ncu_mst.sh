# ncu --set full capture of the MST chain-step kernel (config given as $1, default config2)
mkdir -p gpurun_out
WL=${1:-config2}
CMD="python scripts/dev/quick.py $WL --chains 1024 --S 4 --reps 1"
$CMD > gpurun_out/plain_mst_$WL.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 1 -c 1 -f -o gpurun_out/prof_mst_$WL $CMD > gpurun_out/ncu_mst_$WL.log 2>&1
tail -2 gpurun_out/ncu_mst_$WL.log
