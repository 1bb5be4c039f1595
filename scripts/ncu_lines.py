#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K`.
usage: ... | python scripts/ncu_lines.py [min_pct]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
minpct = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
files = {}
cur = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iL, iS, iA = 0, 1, 2
        iSm = hdr.index("# Samples")
        iEx = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "Function Name":
        continue
    if r[iL]:                      # a CUDA source line row (aggregated)
        try:
            ex = int(r[iEx] or 0); sm = int(r[iSm] or 0)
        except ValueError:
            continue
        files.setdefault(cur, []).append((int(r[iL]), r[iS].strip(), ex, sm))
tot_ex = sum(x[2] for v in files.values() for x in v) or 1
tot_sm = sum(x[3] for v in files.values() for x in v) or 1
print("total instr %d samples %d" % (tot_ex, tot_sm))
for f, v in files.items():
    fe = sum(x[2] for x in v)
    if fe * 100.0 / tot_ex < 0.3:
        continue
    print("== %s  instr %.1f%%  samples %.1f%%" % (f, fe * 100.0 / tot_ex, sum(x[3] for x in v) * 100.0 / tot_sm))
    for ln, src, ex, sm in v:
        if ex * 100.0 / tot_ex >= minpct or sm * 100.0 / tot_sm >= minpct:
            print("  %4d  i %5.1f%%  s %5.1f%%  %s" % (ln, ex * 100.0 / tot_ex, sm * 100.0 / tot_sm, src[:100]))
