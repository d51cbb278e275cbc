#!/bin/bash
# A/B timing of library variants (built with irsgmcmc_b200/build.py `defines=..., out=...`): SVF tests on the default
# build first, then the op microbench and the bench stage times per variant.  Usage: tools/gpu_ab.sh lib1.so lib2.so ...
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sgld.py -m gpu -x -q > gpurun_out/t_ab.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/t_ab.log
tail -3 gpurun_out/t_ab.log
for lib in "$@"; do
  echo "=== $lib"
  IRSGMCMC_LIB=$PWD/irsgmcmc_b200/$lib timeout 300 python tools/microbench.py --size 128 2>&1 | grep -E "svf|langevin|warp3d " | tee gpurun_out/mb_$lib.log
  IRSGMCMC_LIB=$PWD/irsgmcmc_b200/$lib timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 10 --e2e-steps 5 2>/dev/null | tail -1 > gpurun_out/bench_$lib.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$lib.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'])
print({k: round(v,4) for k,v in d.get('stage_ms',{}).items()})
PY
done
