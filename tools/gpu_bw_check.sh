for lib in libirsgmcmc.so libirsgmcmc_bw64.so; do for r in 0 1; do
IRSGMCMC_LIB=$PWD/irsgmcmc_b200/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 2 --as-rank $r 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib as-rank $r', 'ms', round(d['ms_per_step'],4), 'fwd', d['stage_ms']['svf_fwd'])"
done; done
