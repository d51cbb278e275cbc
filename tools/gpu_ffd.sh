#!/bin/bash
# FFD kernels and the SVFFD transition: parity tests, then the FFD timings at 128^3
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_ffd.py tests/test_golden.py -m gpu -q --tb=short -s -k "ffd" > gpurun_out/t_ffd.log 2>&1; rc=$?; echo "ffd tests exit $rc" >> gpurun_out/t_ffd.log
tail -40 gpurun_out/t_ffd.log
[ $rc -ne 0 ] && exit 1
timeout 60 python tools/microbench.py --ffd 4 > gpurun_out/mb_ffd.log 2>&1; timeout 60 python tools/microbench.py --ffd 2 >> gpurun_out/mb_ffd.log 2>&1
cat gpurun_out/mb_ffd.log
