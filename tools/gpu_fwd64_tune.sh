#!/bin/bash
# forward ring at pitch 64: resident CTAs per SM and z-segment length (run-time development switches), on two draws of the state
run() {
env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 --as-rank $R 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('as-rank $R', '$*', 'ms', round(d['ms_per_step'],4), 'fwd', d['stage_ms']['svf_fwd'])"
}
for R in 0 1; do
run IRS_NOP=1
run IRS_FWD_CTAS=4
run IRS_FWD_CTAS=6
run IRS_SVF_SEG_FWD=6
run IRS_SVF_SEG_FWD=11
done
