#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of the built library (cuobjdump -sass), selected families.
usage: python scripts/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "deepblast_b200", "libb200dp.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
FAM = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "LDGSTS", "SYNCS", "ELECT", "MAPA", "UCGABAR", "CGABAR", "MUFU",
       "SHFL", "FMNMX3", "F2FP", "HMMA", "STG", "LDG", "ATOMG", "STS", "LDS", "ST", "LD"]
print("cuobjdump -sass deepblast_b200/libb200dp.so: instruction counts per kernel (static), selected opcode families")
print("(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor, UTMAPF = its L2 prefetch,")
print(" UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier, MAPA / UCGABAR = cluster address mapping / barrier.cluster,")
print(" ST / LD without a space suffix = generic or shared::cluster accesses; HMMA would be the legacy mma.sync path: none)\n")
cur, cnt, order = None, {}, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        cnt[cur]["total"] += 1
        for f in FAM:
            if op == f or op.startswith(f + "."):
                cnt[cur][f] += 1
                break
        else:
            for f in FAM:
                if op.startswith(f):
                    cnt[cur][f] += 1
                    break
names = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
for mangled, name in sorted(zip(order, names), key=lambda x: x[1]):
    c = cnt[mangled]
    short = re.sub(r"^void b200dp::|^b200dp::", "", name)
    short = re.sub(r"\(.*$", "", short)
    print(short)
    print("   total %d  " % c["total"] + "  ".join("%s:%d" % (f, c[f]) for f in FAM if c[f]))
