# usage: bash scripts/gpu_bench_prof.sh <tag> [extra bench args]
tag=$1; shift
mkdir -p gpurun_out
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --no-cpu-baseline --workload config3 "$@" > gpurun_out/bench_${tag}_c3.json 2>> gpurun_out/bench_$tag.err
CMD="python bench.py --steps 2 --warmup 1 --chain-steps 2 --no-cpu-baseline $@"
$CMD > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:recom_step -s 3 -c 1 -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
