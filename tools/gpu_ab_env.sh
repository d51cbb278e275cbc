#!/bin/bash
# A/B of a run-time switch: tools/gpu_ab_env.sh VAR  (runs the bench with VAR=1 and VAR=0)
for v in 1 0; do
env $1=$v python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1=$v', 'ms', round(d['ms_per_step'],4), 'G', round(d['value']/1e9,3), 'launches', d['gpu_launches']//100, {k: round(x,4) for k,x in d['stage_ms'].items()})"
done
