#!/usr/bin/env python
"""Barrier-delimited segments and opcode mixes of a kernel from `ncu --page source --csv --print-source sass`.
usage: python tools/ncu_segments.py sass.csv <units> [seg_index ...]   (units = pencils x panels, to normalise)"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
head = rows[h]; ci = {n: i for i, n in enumerate(head)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(head):
        continue
    try:
        a = int(r[0], 16) if r[0].startswith('0x') else int(r[0])
    except ValueError:
        continue
    data.append((a, r[1].strip(), int(r[ci['Instructions Executed']] or 0), int(r[ci['# Samples']] or 0),
                 int(r[ci['L1 Wavefronts Shared']] or 0)))
base = data[0][0]
tot = sum(d[2] for d in data)
segs, cur = [], []
for d in data:
    cur.append(d)
    if 'BAR.' in d[1] or 'EXIT' in d[1] or d[1].startswith('RET'):
        segs.append(cur); cur = []
if cur:
    segs.append(cur)
want = [int(x) for x in sys.argv[3:]]
for i, s in enumerate(segs):
    n = sum(d[2] for d in s)
    if n / tot > 0.003 or i in want:
        print(f"seg {i:3d} {s[0][0]-base:#7x}-{s[-1][0]-base:#7x} inst {n/tot*100:5.1f}% ({n/units:7.1f}/unit) samples {sum(d[3] for d in s):6d} "
              f"smem wf {sum(d[4] for d in s)/units:7.1f}/unit  ends {s[-1][1][:40]}")
    if i in want:
        ops = collections.Counter()
        for a, src, n2, sm, wf in s:
            t = src.split()
            op = t[1] if t[0].startswith('@') else t[0]
            ops[op.split('.')[0]] += n2
        print("      ", {k: round(v / units, 1) for k, v in ops.most_common(24)})
print('total inst/unit', round(tot / units, 1))
