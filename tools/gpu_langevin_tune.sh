for sz in "256 1" "128 16" "64 64"; do set -- $sz
for cfg in "0 512 32" "1 512 32" "1 256 16"; do set -- $sz $cfg
IRS_LANGEVIN_FUSED=$3 IRS_LANGEVIN_T=$4 IRS_LANGEVIN_TY=$5 python bench.py --size $1 --chains $2 --steps 10 --warmup 3 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1^3 x $2 fused=$3 T=$4 TY=$5', 'ms', round(d['ms_per_step'],4), 'langevin+sobolev', d['stage_ms']['langevin+sobolev'])"
done; done
