#!/usr/bin/env python
"""Joins the stall samples of one kernel in an .ncu-rep (`--import-source on`) with source lines.

    tools/ncu_source_hotspots.py <report.ncu-rep> <kernel-substring> <cubin-disassembly-with-lineinfo> [top]

The disassembly is `nvdisasm -c -g` of the cubin extracted from libasac_b200.so
(`cuobjdump -xelf all libasac_b200.so`).  Prints samples per device function and per source line
with the dominant stall reasons."""
import collections
import csv
import re
import subprocess
import sys

rep, kern, sass = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
line_of, in_kernel, sub, cur = {}, False, 'body', (None, 0)
for l in open(sass):
    m = re.match(r'^\.text\.(\S+):', l)
    if m:
        in_kernel, sub = kern in m.group(1), 'body'
        continue
    m = re.match(r'^\s+\.type\s+(\$\S+),@function', l)
    if m:
        sub = m.group(1).split('$')[-1][:44]
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and in_kernel:
        line_of[int(m.group(1), 16)] = (sub,) + cur
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iA, iS, iI = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base, tot, sec = None, 0, 0
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
by_fn = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        sec += 1
        continue
    if sec > 0:
        break
    if len(r) <= iI or r[iA] == 'Address':
        continue
    a = int(r[iA], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ('?', '?', 0))
    n = int(r[iS])
    tot += n
    for d, k in ((by_line, key), (by_fn, key[0])):
        d[k][0] += n
        d[k][1] += int(r[iI])
        for i in stall:
            v = int(r[i]) if r[i] else 0
            if v:
                d[k][2][hdr[i][6:]] += v
print(f'{kern}: {tot} stall samples (first profiled launch)')
for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f'  {k:46s} {v[0]:6d} ({v[0] / tot:5.1%}) inst={v[1]:8d}  {dict(v[2].most_common(4))}')
print()
for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'  {k[0][:28]:28s} {k[1]}:{k[2]:<5d} {v[0]:6d} ({v[0] / tot:5.1%}) inst={v[1]:8d}  {dict(v[2].most_common(3))}')
