for r in 0 1 5; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --as-rank $r 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('as-rank $r', 'ms', round(d['ms_per_step'],4), 'clocks', d['clocks'], {k: d['stage_ms'][k] for k in ('svf_fwd','svf_adjoint','warp','mixture_step')}, d['svf_max_abs_u_per_step'][-4:])"
done
