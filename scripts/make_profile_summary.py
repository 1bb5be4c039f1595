#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, prof_fwd_bwd.ncu-rep} into the tracked summaries under
profiles/ (round-tagged) and profiles/traffic.json (per-launch DRAM bytes used by bench.py).
usage: python scripts/make_profile_summary.py r01 [workload]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list ----------------------------------------------------------------------
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
launches = [(r[h.index("Kernel Name")], float(r[h.index("Metric Value")])) for r in rows[hi + 1:] if len(r) == len(h)]
with open(os.path.join(P, f"{tag}_launches_{wl}.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 60 python bench.py --steps 2 --warmup 1 "
            "--no-cpu-baseline --no-e2e   (per-launch times are cold-cache and serialised)\n")
    f.write("kernel,duration_ns\n")
    for k, d in launches:
        f.write('"%s",%d\n' % (k.replace('"', "'"), d))
ours = [(k, d) for k, d in launches if "softdp" in k]
# the step's own launches: skip input generation and the untimed spin kernel bench.py queues
steps = [(k, d) for k, d in launches if not any(x in k for x in ("distribution_elementwise", "neg_kernel", "spin_kernel"))]
tot = sum(d for _, d in steps) or 1.0
share = {}
for k, d in steps:
    key = "softdp_fwd" if "softdp_fwd" in k else "softdp_bwd" if "softdp_bwd" in k else "other (torch sum/fill)"
    share[key] = share.get(key, 0.0) + d
# ---- full capture -----------------------------------------------------------------------
rep = os.path.join(G, "prof_fwd_bwd.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
tp = os.path.join(P, "traffic.json")
if os.path.exists(tp):
    traffic = json.load(open(tp))
with open(os.path.join(P, f"{tag}_ncu_summary_{wl}.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:softdp_ -s 6 -c 2 python bench.py "
            "--steps 2 --warmup 1 --no-cpu-baseline --no-e2e\n(workload %s; units as printed by ncu: time us, bytes MB)\n\n" % wl)
    f.write("share of the step (launch list, steps only): " +
            ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(share.items())) + "\n\n")
    for r in rr[2:]:
        name = r[hdr.index("Kernel Name")]
        f.write("== %s\n" % name)
        vals = {}
        for w in want:
            if w in hdr:
                vals[w] = r[hdr.index(w)]
                f.write("  %-62s %s\n" % (w, r[hdr.index(w)]))
        try:
            tb = (float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])) * 1e6
            traffic["%s_%s" % (wl, "fwd" if "fwd" in name else "bwd")] = tb
        except Exception:
            pass
        kn = "softdp_fwd" if "fwd" in name else "softdp_bwd"
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kn],
                             capture_output=True, text=True).stdout
        top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_top.py"), "12"], input=src,
                             capture_output=True, text=True).stdout
        f.write("\n".join("  " + l for l in top.splitlines()[1:]) + "\n\n")
json.dump(traffic, open(tp, "w"), indent=1)
print(open(os.path.join(P, f"{tag}_ncu_summary_{wl}.txt")).read()[:3000])
