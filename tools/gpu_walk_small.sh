for w in 1 0; do for c in 1 2 4; do IRS_GMM_WALK=$w python bench.py --size 64 --chains $c --data lcc --steps 20 --warmup 3 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('walk=$w chains=$c', round(d['ms_per_step'],4), d['stage_ms']['mixture_step'])"; done; done
