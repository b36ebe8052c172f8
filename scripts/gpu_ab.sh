# A/B of library builds on ONE box: bash scripts/gpu_ab.sh tagA tagB ... (variants/lib_<tag>.so)
mkdir -p gpurun_out
for rep in 1 2; do
for tag in "$@"; do
  for wl in config2 config3; do
    RECOM_B200_LIB=$PWD/variants/lib_$tag.so python bench.py --no-cpu-baseline --workload $wl --steps 40 --warmup 3 2>> gpurun_out/ab.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$tag', '$wl', 'rep$rep', round(d['value']), 'e2e', round(d['e2e']['value']), 'ctas', d['config']['ctas_per_sm'], 'smem', d['config']['smem_bytes_per_cta'])
" >> gpurun_out/ab.log
  done
done
done
