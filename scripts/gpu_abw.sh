# Wilson path A/B on ONE box: variants/lib_<tag>.so, parity tests first
mkdir -p gpurun_out
: > gpurun_out/abw.log
T=$PWD/variants/lib_$2.so
RECOM_B200_LIB=$T timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wilson or reversible" > gpurun_out/abw_tests_$2.log 2>&1
echo "$2 tests rc=$?" >> gpurun_out/abw.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --wilson --warmup 3 $BARGS 2>> gpurun_out/abw.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$tag', '$BARGS', round(d['value']), 'e2e', round(d['e2e']['value']), 'ctas', d['config']['ctas_per_sm'], 'thr', d['config']['threads_per_cta'], 'smem', d['config']['smem_bytes_per_cta'])
" >> gpurun_out/abw.log
}
for tag in "$@"; do
BARGS="--workload config3 --chain-steps 128 --steps 3"
run $tag RECOM_B200_LIB=$PWD/variants/lib_$tag.so
BARGS="--workload config1 --chain-steps 128 --steps 5"
run $tag RECOM_B200_LIB=$PWD/variants/lib_$tag.so
BARGS="--workload config4 --chain-steps 8 --steps 2"
run $tag RECOM_B200_LIB=$PWD/variants/lib_$tag.so
done
