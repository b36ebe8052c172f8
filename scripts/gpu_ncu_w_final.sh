# round 2, last build: ncu --set full of the Wilson chain-step kernel at config 3 after the 16-bit adjacency
# (5 CTAs of 192 threads per SM); the summary replaces the earlier capture under profiles/.
mkdir -p gpurun_out /tmp/rep
CMD="python scripts/dev/quick.py config3 --wilson --chains 1024 --S 16 --reps 2"
RECOM_B200_SHARE=0 timeout 100 $CMD > gpurun_out/r2_plain_wilson_final.log 2>&1 && \
RECOM_B200_SHARE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:recom_step -s 2 -c 1 -f -o /tmp/rep/wf $CMD > gpurun_out/r2_ncu_wilson_final.log 2>&1 && \
python scripts/summarize_ncu.py /tmp/rep/wf.ncu-rep "rc::recom_step_kernel<Wilson> on config 3 (Delaunay 9000 nodes, 18 districts, eps 0.02), 1024 chains x 16 steps, 4 walkers, 16-bit adjacency entries, plain build (RECOM_B200_SHARE=0)" 16384 > gpurun_out/r2_recom_step_kernel_wilson_config3_final_ncu.txt 2> gpurun_out/r2_sum_wf.err
tail -1 gpurun_out/r2_plain_wilson_final.log; head -12 gpurun_out/r2_recom_step_kernel_wilson_config3_final_ncu.txt
