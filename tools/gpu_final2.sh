#!/bin/bash
# last evidence run of round 2 on one B200 (the captures of tools/gpu_round2.sh stay valid for the SVF / mixture / VI kernels)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; echo "exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
bash tools/gpu_launchlist.sh
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 1"
ncu --set full --clock-control none --import-source on -k "regex:langevin_xy|smooth_axis|warp_vox_fwd|box_march|sgd_update" -s 30 -c 8 \
    -o gpurun_out/prof_small -f $B > gpurun_out/ncu_small.log 2>&1
bash tools/gpu_ffd_ncu.sh > gpurun_out/ffd_kernels.txt 2>&1
ls -la gpurun_out/*.ncu-rep
