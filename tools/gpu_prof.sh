#!/bin/bash
# ncu full capture of the kernels matching $1 (regex), skipping $2 launches, $3 launches captured -> gpurun_out/$4.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${2:-0} -c ${3:-6} -o gpurun_out/${4:-prof} -f \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
