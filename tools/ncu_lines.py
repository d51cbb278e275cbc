"""Per-source-line executed instructions / stall samples of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py rep kernel_regex [launch_skip] [min_pct]"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, kern, skip='0', min_pct='0.7'):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                          'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        elif len(r) > 20 and r[0] == 'Line No':
            hdr = r
            ie, ss = hdr.index('Instructions Executed'), hdr.index('# Samples')
        elif len(r) > 20 and hdr:
            if r[0] != '':
                last = (cur_file, int(r[0]), r[1].strip()[:110])
            key = last
            a = agg.setdefault(key, [0, 0, 0])
            try:
                a[0] += int(r[ie]); a[1] += int(r[ss]); a[2] += 1
            except ValueError:
                pass
    ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print(f'# total warp instructions {ti}, samples {ts}')
    for (f, ln, src), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if a[0] * 100.0 / max(ti, 1) >= float(min_pct) or a[1] * 100.0 / max(ts, 1) >= float(min_pct):
            print(f'{f:16s}{ln:5d} inst {100.0 * a[0] / ti:5.1f}% samp {100.0 * a[1] / max(ts,1):5.1f}% sass {a[2]:4d} | {src}')


if __name__ == '__main__':
    main(*sys.argv[1:])
