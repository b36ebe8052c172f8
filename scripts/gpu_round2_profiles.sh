# round 2: ncu --set full captures of the chain-step kernel on HEAD (MST config 2 / config 3, Wilson config 3)
# + the launch list of the default bench command.  Each ncu run only after the same command exited 0 without it.
mkdir -p gpurun_out
prof() {  # tag, quick.py args
  TAG=$1; shift
  CMD="python scripts/dev/quick.py $*"
  $CMD > gpurun_out/r2_plain_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:recom_step -s 1 -c 1 -f -o gpurun_out/r2_prof_$TAG $CMD > gpurun_out/r2_ncu_$TAG.log 2>&1
  tail -1 gpurun_out/r2_plain_$TAG.log
}
prof mst_c2 config2 --chains 1024 --S 32 --reps 1
prof mst_c3 config3 --chains 1024 --S 32 --reps 1
prof wilson_c3 config3 --wilson --chains 1024 --S 8 --reps 1
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras config3,config3_wilson"
$BENCH > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_bench_ncu.log 2>&1
echo done
