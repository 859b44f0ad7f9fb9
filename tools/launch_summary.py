#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals over the LAST full iteration."""
import csv, collections, re, sys
def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(row['Metric Value'].replace(',', ''))
            u = row['Metric Unit']
            v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
            rows.append((row['Kernel Name'], v, row['Grid Size'], row['Block Size']))
    return rows
def last_iteration(rows, marker='tanh_bwd'):
    idx = [i for i, r in enumerate(rows) if marker in r[0]]
    return rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
if __name__ == '__main__':
    rows = load(sys.argv[1])
    seg = last_iteration(rows)
    if len(sys.argv) > 2 and sys.argv[2] == 'seq':
        for nm, v, g, b in seg: print(f'{v:8.1f} us  {g:>16s} {b:>14s}  {re.sub(r"\(.*", "", nm)[:90]}')
    agg = collections.OrderedDict()
    for nm, v, g, b in seg:
        k = re.sub(r'\(.*', '', nm); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in agg.values())
    print(f'kernels/iter {len(seg)}  sum {tot:.1f} us')
    for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]): print(f'  {v:8.1f} us {c:3d}  {k[:110]}')
