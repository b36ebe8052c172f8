mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_29.log
rm -f gpurun_out/thr.log
for wl in config1 config2 config3 config5; do
python bench.py --no-cpu-baseline --workload $wl --steps 12 --warmup 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl', 'auto threads', d['config']['threads_per_cta'], round(d['value']), 'e2e', round(d['e2e']['value']), 'ctas/sm', d['config']['ctas_per_sm'], 'smem', d['config']['smem_bytes_per_cta'])" >> gpurun_out/thr.log
done
