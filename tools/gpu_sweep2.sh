#!/bin/bash
# sweep an environment variable and print the bench value and all stage times
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$VAR=$v', round(d['value'] / 1e9, 4), 'G', round(d['ms_per_step'], 4), 'ms', {k: round(x, 3) for k, x in d['stage_ms'].items()})
"
done
