#!/bin/bash
# ncu evidence: launch list of the bench command + one full capture of the two hot kernels.
mkdir -p gpurun_out
echo "== timing"; timeout 300 python -u scripts/gpu_time.py 2>&1 | tee gpurun_out/time_sweep.log
echo "== pytest parity"; timeout 900 python -u -m pytest tests/test_gpu_parity.py -m gpu -q -x --maxfail=3 > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/pytest_parity.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?"; grep -c softdp gpurun_out/launches.csv
echo "== ncu full (fwd, bwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softdp_ -s 6 -c 2 -f -o gpurun_out/prof_fwd_bwd \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
