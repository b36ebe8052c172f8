# ncu --set full of the shared-retry build (RECOM_B200_SHARE=1) on config 3, same launch as r2_prof_mst_c3
mkdir -p gpurun_out
CMD="python scripts/dev/quick.py config3 --chains 1024 --S 64 --reps 2"
RECOM_B200_SHARE=1 timeout 120 $CMD > gpurun_out/r2_share_c3.log 2>&1 && \
RECOM_B200_SHARE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:recom_step -s 2 -c 1 -f -o gpurun_out/r2_prof_mst_c3_share $CMD > gpurun_out/r2_ncu_share.log 2>&1
tail -1 gpurun_out/r2_share_c3.log
