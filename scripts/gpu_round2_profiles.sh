# round 2: ncu --set full captures of the chain-step kernel on HEAD (MST config 2 / config 3, Wilson config 3)
# + the launch list of a short bench command.  Each ncu run only after the same command exited 0 without it.
mkdir -p gpurun_out
prof() {  # tag, quick.py args
  TAG=$1; shift
  CMD="python scripts/dev/quick.py $*"
  RECOM_B200_SHARE=0 timeout 120 $CMD > gpurun_out/r2_plain_$TAG.log 2>&1 && \
  RECOM_B200_SHARE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:recom_step -s 2 -c 1 -f -o gpurun_out/r2_prof_$TAG $CMD > gpurun_out/r2_ncu_$TAG.log 2>&1
  tail -1 gpurun_out/r2_plain_$TAG.log
}
# (launch 0: the 4 warm-up steps; launch 1: rep 1; launch 2 is profiled: chains past the hard pairs of the seed plan)
prof mst_c2 config2 --chains 1024 --S 64 --reps 2
prof mst_c3 config3 --chains 1024 --S 64 --reps 2
prof wilson_c3 config3 --wilson --chains 1024 --S 16 --reps 2
# the shared-retry build on the same config-3 launch
CMD="python scripts/dev/quick.py config3 --chains 1024 --S 64 --reps 2"
RECOM_B200_SHARE=1 timeout 120 $CMD > gpurun_out/r2_plain_mst_c3_share.log 2>&1 && \
RECOM_B200_SHARE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:recom_step -s 2 -c 1 -f -o gpurun_out/r2_prof_mst_c3_share $CMD > gpurun_out/r2_ncu_mst_c3_share.log 2>&1
tail -1 gpurun_out/r2_plain_mst_c3_share.log
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras config3,config3_wilson"
timeout 200 $BENCH > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_bench_ncu.log 2>&1
echo done
