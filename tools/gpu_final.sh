#!/bin/bash
# one GPU-box visit at HEAD: all GPU tests, smoke, FFD timings, the secondary SVFFD bench line and the FFD kernels' ncu durations
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q --tb=short > gpurun_out/t_all.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/t_all.log
tail -25 gpurun_out/t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 60 python tools/microbench.py --ffd 4 > gpurun_out/mb_ffd.log 2>&1; cat gpurun_out/mb_ffd.log
timeout 120 python bench.py --cps 4 --steps 100 --warmup 5 --e2e-steps 20 --no-cpu-baseline > gpurun_out/bench_svffd4.log 2>&1; echo "exit $?" >> gpurun_out/bench_svffd4.log
tail -2 gpurun_out/bench_svffd4.log | cut -c1-2500
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ffd_axis -s 120 -c 12 --csv --log-file gpurun_out/launches_ffd.csv \
    python bench.py --cps 4 --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_ffd.log 2>&1
grep -c ffd_axis gpurun_out/launches_ffd.csv; tail -12 gpurun_out/launches_ffd.csv | cut -c1-400
