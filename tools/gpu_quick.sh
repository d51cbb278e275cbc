#!/bin/bash
# quick GPU check: SVF op tests first (stop on failure / hang), then everything, then a short bench
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k svf > gpurun_out/t_svf.log 2>&1; rc=$?; echo "svf tests exit $rc" >> gpurun_out/t_svf.log
tail -15 gpurun_out/t_svf.log
[ $rc -ne 0 ] && exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
[ $rc -ne 0 ] && exit 1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "exit $?" >> gpurun_out/bench_quick.log
tail -2 gpurun_out/bench_quick.log | cut -c1-3000
