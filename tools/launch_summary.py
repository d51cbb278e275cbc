"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / time / share of ONE transition
(the last complete one in the list).  Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('<unnamed>::', '').replace('void ', '') for r in rows]
    t = [float(r['Metric Value']) / 1e3 for r in rows]
    idx = [i for i, n in enumerate(names) if n.startswith('langevin')]
    print(f'# {len(rows)} launches in the list; transitions start at launch ids {idx}')
    if len(idx) < 2:
        s, e = (idx[0] if idx else 0), len(rows)
    else:   # the shortest slice = a device-resident transition (the end-to-end steps add the image (re)load kernels)
        s, e = min(zip(idx[1:-1], idx[2:]), key=lambda p: p[1] - p[0]) if len(idx) > 2 else (idx[-2], idx[-1])
    agg = collections.OrderedDict()
    for n, x in zip(names[s:e], t[s:e]):
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += x
    tot = sum(v[1] for v in agg.values())
    print(f'# one transition = launches [{s}, {e}): {e - s} launches, {tot:.1f} us summed (cold-cache, serialised)')
    for n, v in agg.items():
        print(f'{n:60s} n={v[0]:3d}  {v[1]:9.1f} us  {v[1] / v[0]:8.1f} us/launch  {100 * v[1] / tot:5.1f} %')


if __name__ == '__main__':
    main(sys.argv[1])
