#!/usr/bin/env python
"""Aggregates `ncu --page source --csv --print-source sass,cuda` output per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_hot.py src.csv [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
STALLS = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg",
          "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_not_selected", "stall_selected",
          "stall_membar", "stall_sleep", "stall_tex", "stall_drain", "stall_misc")
tot_inst = tot_samp = 0
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
allst = collections.Counter()
for si, h in enumerate(hdr_idx):
    head = rows[h]; fp = rows[h - 2][1] if h >= 2 else "?"
    end = hdr_idx[si + 1] - 2 if si + 1 < len(hdr_idx) else len(rows)
    ci = {n: i for i, n in enumerate(head)}
    for r in rows[h + 1:end]:
        if len(r) < len(head):
            continue
        try:
            ln = int(r[0]); inst = int(r[ci["Instructions Executed"]] or 0); samp = int(r[ci["# Samples"]] or 0)
        except ValueError:
            continue
        key = (fp.split("/")[-1], ln, r[1].strip()[:100])
        agg[key][0] += inst; agg[key][1] += samp
        for st in STALLS:
            v = r[ci[st]] if st in ci else ""
            if v:
                agg[key][2][st] += int(v); allst[st] += int(v)
        tot_inst += inst; tot_samp += samp
print("total warp instructions", tot_inst, "samples", tot_samp)
print("stall mix:", {k: f"{v / max(tot_samp, 1) * 100:.1f}%" for k, v in allst.most_common(8)})
for (f, ln, src), (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{f}:{ln:4d} inst {inst / tot_inst * 100:5.1f}% samp {samp / tot_samp * 100:5.1f}% "
          f"{dict(st.most_common(2))} | {src}")
