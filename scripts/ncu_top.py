#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest
SASS instructions.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | python scripts/ncu_top.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] in ("Kernel Name", "Address")), len(rows))
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
all_s = sum(tot.values()) or 1
print("kernel:", rows[0][1] if rows and len(rows[0]) > 1 else "?")
print("stall totals (%% of %d samples):" % all_s)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-26s %6.2f%%" % (h, 100.0 * v / all_s))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
print("instructions executed (warp-level) total: %d" % sum(int(r[col["Instructions Executed"]] or 0) for r in body))
print("top %d instructions by samples:" % n)
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[col[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("  %5d %6s  %-60s %s" % (i, r[col["# Samples"]], r[col["Source"]].strip()[:60],
                                   " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
