#!/bin/bash
# Round-end evidence: the driver's bench line, the reference arm, the ncu launch list of the same command
# (times only, and once more with DRAM / occupancy metrics), full captures of the hot kernels.
T=${TAG:-r02f}
mkdir -p gpurun_out
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/${T}_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -c 600 --csv --log-file gpurun_out/${T}_traffic.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${T}_under_ncu2.log 2>&1; echo "ncu traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'softdp_fwd3|softdp_sq_bwd' -s 6 -c 2 -f -o gpurun_out/${T}_c2 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras --no-graph > gpurun_out/${T}_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'softdp_sq_fwd|softdp_sq_bwd' -c 2 -f -o gpurun_out/${T}_b32 \
    python scripts/gpu_sq_one.py b32 0 1 > gpurun_out/${T}_ncu_b32.log 2>&1; echo "ncu b32 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'softdp_cl_fwd|softdp_traceback' -c 4 -f -o gpurun_out/${T}_small \
    python scripts/gpu_small_one.py > gpurun_out/${T}_ncu_small.log 2>&1; echo "ncu small rc=$?"
ls -la gpurun_out/${T}_*
