# round 2: ncu --set full captures of the chain-step kernel on HEAD (MST config 2 / config 3, plain and
# shared-retry build, Wilson config 3) + the launch list of a short bench command.  Each ncu run only
# after the same command exited 0 without it.  The reports are summarised ON THE BOX (scripts/
# summarize_ncu.py) and deleted: four of them exceed what gpurun copies back.
mkdir -p gpurun_out /tmp/rep
prof() {  # tag, share mode, steps in the launch, title, quick.py args
  TAG=$1; SH=$2; STEPS=$3; TITLE=$4; shift 4
  CMD="python scripts/dev/quick.py $*"
  RECOM_B200_SHARE=$SH timeout 120 $CMD > gpurun_out/r2_plain_$TAG.log 2>&1 && \
  RECOM_B200_SHARE=$SH timeout 600 ncu --set full --clock-control none --import-source on -k regex:recom_step -s 2 -c 1 -f -o /tmp/rep/$TAG $CMD > gpurun_out/r2_ncu_$TAG.log 2>&1 && \
  python scripts/summarize_ncu.py /tmp/rep/$TAG.ncu-rep "$TITLE" $STEPS > gpurun_out/r2_recom_step_kernel_${TAG}_ncu.txt 2> gpurun_out/r2_sum_$TAG.err
  tail -1 gpurun_out/r2_plain_$TAG.log
}
# (launch 0: the 4 warm-up steps; launch 1: rep 1; launch 2 is profiled: chains past the hard pairs of the seed plan)
prof mst_config2 0 65536 "rc::recom_step_kernel<MST> on config 2 (100x100 grid, 10 districts, eps 0.05), 1024 chains x 64 steps, plain build (RECOM_B200_SHARE=0)" config2 --chains 1024 --S 64 --reps 2
prof mst_config3 0 65536 "rc::recom_step_kernel<MST> on config 3 (Delaunay 9000 nodes, 18 districts, eps 0.02), 1024 chains x 64 steps, plain build (RECOM_B200_SHARE=0)" config3 --chains 1024 --S 64 --reps 2
prof mst_config3_share 1 65536 "rc::recom_step_kernel<MST, shared retries> on config 3, 1024 chains x 64 steps (RECOM_B200_SHARE=1)" config3 --chains 1024 --S 64 --reps 2
prof wilson_config3 0 16384 "rc::recom_step_kernel<Wilson> on config 3 (Delaunay 9000 nodes, 18 districts, eps 0.02), 1024 chains x 16 steps, 4 walkers, plain build (RECOM_B200_SHARE=0)" config3 --wilson --chains 1024 --S 16 --reps 2
cp /tmp/rep/mst_config2.ncu-rep gpurun_out/r2_prof_mst_c2.ncu-rep 2>/dev/null
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras config3,config3_wilson"
timeout 200 $BENCH > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_bench_ncu.log 2>&1
echo done
