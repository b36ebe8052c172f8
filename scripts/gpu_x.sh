mkdir -p gpurun_out
for cs in 16 32 64 128; do
python bench.py --no-cpu-baseline --chain-steps $cs --steps $((3200/cs)) --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('chain_steps', $cs, round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['samples'])" >> gpurun_out/cs.log
done
