# A/B of library builds on ONE box: bash scripts/gpu_ab.sh tagA tagB ... (variants/lib_<tag>.so);
# the parity tests run first on the second tag
mkdir -p gpurun_out
: > gpurun_out/ab.log
RECOM_B200_LIB=$PWD/variants/lib_$2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/ab_tests_$2.log 2>&1
echo "$2 tests rc=$?" >> gpurun_out/ab.log
for rep in 1; do
for tag in "$@"; do
  for wl in config2 config3; do
    RECOM_B200_LIB=$PWD/variants/lib_$tag.so timeout 300 python bench.py --no-cpu-baseline --workload $wl --steps 15 --warmup 3 2>> gpurun_out/ab.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$tag', '$wl', 'rep$rep', round(d['value']), 'e2e', round(d['e2e']['value']), 'ctas', d['config']['ctas_per_sm'], 'thr', d['config']['threads_per_cta'], 'smem', d['config']['smem_bytes_per_cta'])
" >> gpurun_out/ab.log
  done
done
done
