#!/bin/bash
# Turn the outputs of scripts/gpu_final.sh (gpurun_out/${T}_*) into the tracked files under profiles/.
T=${1:-r02g}; R=${2:-r02}
cp gpurun_out/${T}_bench_n1.json profiles/${R}_bench_n1.json
cp gpurun_out/${T}_ref.json profiles/${R}_bench_reference_arm.json
( echo "# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline   (per-launch times are cold-cache and serialised)"
  grep -v "^==" gpurun_out/${T}_launches.csv | python -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[hi]
print('kernel,grid,duration_ns')
for r in rows[hi+1:]:
    if len(r)==len(h):
        n=r[h.index('Kernel Name')]
        m=re.match(r'void b200dp::(\w+)<(.*)>\(',n) or re.match(r'void b200dp::(\w+)\(',n) or re.match(r'b200dp::(\w+)\(',n)
        short=(m.group(1)+('<'+m.group(2)+'>' if m.lastindex and m.lastindex>1 else '')) if m else n[:50]
        print('\"%s\",\"%s\",%s'%(short.replace('(bool)','').replace('(int)',''), r[h.index('Grid Size')], r[h.index('Metric Value')]))
" ) > profiles/${R}_launches_bench.csv
python scripts/ncu_launch_table.py gpurun_out/${T}_traffic.csv > /tmp/${T}_table.txt 2>&1
( echo "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active...,smsp__issue_active...,launch__registers_per_thread --clock-control none python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-e2e"
  echo "(one line per kernel instantiation and grid: launches, mean duration, mean DRAM bytes per launch; cold cache, serialised, several passes per kernel)"; echo
  grep -E "softdp|kernel<" /tmp/${T}_table.txt ) > profiles/${R}_launch_table_bench.txt
python scripts/ncu_report_summary.py gpurun_out/${T}_c2.ncu-rep profiles/${R}_ncu_summary_c2.txt "ncu --set full --clock-control none --import-source on -k regex:'softdp_fwd3|softdp_sq_bwd' -s 6 -c 2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras --no-graph   (BASELINE configs[1], the two kernels of the headline step)"
python scripts/ncu_report_summary.py gpurun_out/${T}_b32.ncu-rep profiles/${R}_ncu_summary_b32.txt "ncu --set full --clock-control none --import-source on -k regex:'softdp_sq_fwd|softdp_sq_bwd' -c 2 python scripts/gpu_sq_one.py b32 0 1   (32 pairs 1024 x 1024: forward with TMA operand staging, backward; chain-bound)"
python scripts/ncu_report_summary.py gpurun_out/${T}_small.ncu-rep profiles/${R}_ncu_summary_small_batch.txt "ncu --set full --clock-control none --import-source on -k regex:'softdp_cl_fwd|softdp_traceback' -c 4 python scripts/gpu_small_one.py   (64 pairs 256 x 256 through Decoder.decode + traceback_batch: cluster forward, batched walk)"
