#!/bin/bash
# closing evidence run of round 2 after the forward ring went to pitch 64: whole GPU suite, smoke, and one ncu capture of the forward kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 1"
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:svf_step_fwd_tma" -s 36 -c 4 \
    -o gpurun_out/prof_fwd64 -f $B > gpurun_out/ncu_fwd64.log 2>&1
ls -la gpurun_out/*.ncu-rep
