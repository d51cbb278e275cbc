for st in 20 100 20; do
python bench.py --steps $st --warmup 5 --no-cpu-baseline --no-aten-gpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('steps', d['steps'], 'value', round(d['value']/1e9,3), 'e2e', d['e2e'])"
done
