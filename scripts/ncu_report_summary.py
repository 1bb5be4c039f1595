#!/usr/bin/env python
"""Summarise every kernel of an `ncu --set full` report into a tracked text file under profiles/:
key metrics from the raw page plus stall-reason totals and the hottest SASS instructions (source page).
usage: python scripts/ncu_report_summary.py gpurun_out/X.ncu-rep profiles/r02_ncu_summary_X.txt "command line that produced it" """
import csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
seen = set()
with open(out, "w") as f:
    f.write(cmd + "\n(units as printed by ncu: time us, bytes MB; per-launch, cold cache, serialised)\n\n")
    for r in rr[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        f.write("== %s\n" % name)
        for w in want:
            cols = [h for h in hdr if h == w or h.endswith("." + w)]
            if cols:
                f.write("  %-70s %s\n" % (w, r[hdr.index(cols[0])]))
        base = name.split("<")[0].split("(")[0].replace("void ", "").split("::")[-1]
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + base],
                             capture_output=True, text=True).stdout
        top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_top.py"), "10"], input=src,
                             capture_output=True, text=True).stdout
        f.write("\n".join("  " + l for l in top.splitlines()[1:]) + "\n\n")
print(open(out).read()[:2500])
