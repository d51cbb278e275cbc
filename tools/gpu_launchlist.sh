#!/bin/bash
# ncu launch list of ONE transition of the bench command (our kernels only: the synthetic pair is generated with ATen kernels first)
K="regex:langevin|smooth_|svf_|warp_|box_march|gmm_|reg_hyper|sgd_update|reg_energy|ssd_residual"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 250 -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 1 > gpurun_out/ncu.log 2>&1
grep -c svf_step gpurun_out/launches.csv
