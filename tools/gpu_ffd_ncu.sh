#!/bin/bash
# per-kernel durations, DRAM bytes, instructions and occupancy of the six FFD launches at 128^3, spacing 4
# (tools/microbench.py --ffd runs 23 forward ops = 69 launches, then 23 adjoint ops: launches 60..83 cover both)
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum \
    --clock-control none -k regex:ffd_ -s 60 -c 24 --csv --log-file gpurun_out/ncu_ffd_kernels.csv \
    python tools/microbench.py --ffd 4 > gpurun_out/ncu_ffd_kernels.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/ncu_ffd_kernels.csv')) if len(r) > 10 and r[0].isdigit()]
out = {}
for r in rows:
    out.setdefault((r[0], r[4].split('(')[0][-40:], r[8]), {})[r[12]] = r[14]
for k, v in out.items():
    print(k, v)
PY
