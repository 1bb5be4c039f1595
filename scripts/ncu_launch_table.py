#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv` log of
bench.py: one line per distinct (kernel instantiation, grid) with launch count, mean duration, mean DRAM bytes,
warps active, issue active, registers.  usage: ncu_launch_table.py LOG.csv [--json OUT]"""
import csv, re, sys, json, collections

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = L.setdefault(r[ii], {'name': r[ki]})
    v = float(r[vi].replace(',', ''))
    v *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'ms': 1e3, 'ns': 1e-3}.get(r[ui], 1.0)
    d[r[mi]] = v


def short(n):
    m = re.match(r'void b200dp::(\w+)<(.*)>\(', n)
    if m:
        return m.group(1) + '<' + m.group(2).replace('(bool)', '').replace('(int)', '') + '>'
    m = re.match(r'void b200dp::(\w+)\(', n)
    return m.group(1) if m else n[:60]


agg = collections.OrderedDict()
for d in L.values():
    key = (short(d['name']), int(d.get('launch__grid_size', 0)))
    a = agg.setdefault(key, {'n': 0, 'us': 0.0, 'dram': 0.0, 'warps': 0.0, 'issue': 0.0, 'regs': 0})
    a['n'] += 1
    a['us'] += d.get('gpu__time_duration.sum', 0)
    a['dram'] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a['warps'] += d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0)
    a['issue'] += d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)
    a['regs'] = int(d.get('launch__registers_per_thread', 0))
print(f"{'kernel<template args>':62s} {'grid':>6s} {'n':>4s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>7s} {'warps%':>6s} {'issue%':>6s} {'regs':>4s}")
out = []
for (k, g), a in agg.items():
    n = a['n']
    us, dram = a['us'] / n, a['dram'] / n
    print(f"{k[:62]:62s} {g:6d} {n:4d} {us:9.1f} {dram / 1e6:9.1f} {dram / us / 1e3 if us else 0:7.0f} {a['warps'] / n:6.1f} {a['issue'] / n:6.1f} {a['regs']:4d}")
    out.append({'kernel': k, 'grid': g, 'launches': n, 'us': us, 'dram_bytes': dram, 'warps_active_pct': a['warps'] / n,
                'issue_active_pct': a['issue'] / n, 'regs': a['regs']})
if '--json' in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)
