# segment-length sweep of the SVF kernels at 128^3 x 1 (development overrides IRS_SVF_SEG_BWD / IRS_SVF_SEG_FWD)
run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'ms', round(d['ms_per_step'],4), 'fwd', d['stage_ms']['svf_fwd'], 'bwd', d['stage_ms']['svf_adjoint'])"; }
run default
for s in 5 6 7 8 9 10 11 16; do IRS_SVF_SEG_BWD=$s run "bwd_seg=$s"; done
for s in 4 5 6 8 11 13 16; do IRS_SVF_SEG_FWD=$s run "fwd_seg=$s"; done
