run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'ms', round(d['ms_per_step'],4), 'fwd', d['stage_ms']['svf_fwd'], 'bwd', d['stage_ms']['svf_adjoint'])"; }
for s in 12 13 14 15; do IRS_SVF_SEG_BWD=$s run "bwd_seg=$s"; done
