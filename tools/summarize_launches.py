#!/usr/bin/env python
"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel mean duration
and share of the step (torch's L2-flush fill kernel excluded)."""
import collections
import csv
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name'].split('(')[0].replace('asac::', '')
    key = f"{name} grid={row['Grid Size']} block={row['Block Size']}"
    agg.setdefault(key, []).append(float(row['Metric Value'].replace(',', '')))
own = {k: v for k, v in agg.items() if 'at::' not in k}
tot = sum(sum(v) for v in own.values())
print(f'{"kernel":70s} {"n":>3s} {"mean us":>9s} {"share":>7s}')
for k, v in agg.items():
    share = f'{sum(v) / tot:7.3f}' if k in own else '      -'
    print(f'{k:70s} {len(v):3d} {sum(v) / len(v) / 1e3:9.2f} {share}')
print(f'sum of own kernels per step: {tot / steps / 1e3:.1f} us  ({len(own)} distinct kernels, '
      f'{sum(len(v) for v in own.values()) // steps} launches per step)')
