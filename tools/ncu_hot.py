#!/usr/bin/env python
"""Per-CUDA-source-line hotspots of an .ncu-rep (needs -lineinfo + --import-source on).  Usage: ncu_hot.py rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
fname = ''
func = ''
data = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        func = r[1]
        continue
    if r[0] == 'Line No':
        hdr = r
        ii = hdr.index('Instructions Executed')
        sa = hdr.index('# Samples')
        cf = hdr.index('L1 Wavefronts Shared Excessive') if 'L1 Wavefronts Shared Excessive' in hdr else None
        continue
    if hdr and r[0].isdigit() and len(r) > sa:
        try:
            data.append((int(r[ii]), int(r[sa]), int(r[cf]) if cf and r[cf].isdigit() else 0, fname, r[0], r[1][:105], func[:40]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
ts = sum(d[1] for d in data) or 1
print(f'total warp-inst {tot}, samples {ts}')
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print(f'{100*d[0]/tot:5.1f}% inst {100*d[1]/ts:5.1f}% samp confl={d[2]:>10d} {d[3]}:{d[4]}: {d[5]}')
