#!/bin/bash
for i in 1 2; do
for nt in 2 1; do
IRS_BWD_NT=$nt python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('NT=$nt', round(d['value'] / 1e9, 3), 'e2e', round(d['e2e']['value'] / 1e9, 3), 'serial', round(d['e2e']['serial_value'] / 1e9, 3))
"
done
done
