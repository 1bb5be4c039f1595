#!/bin/bash
# One GPU-box visit: staged bring-up, parity tests, smoke, a short bench.  Every stage has
# its own `timeout` so a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== staged"; timeout 600 python -u scripts/gpu_stage.py v2 2>&1 | tee gpurun_out/stage.log
echo "== timing"; timeout 300 python -u scripts/gpu_time.py 2>&1 | tee gpurun_out/time_sweep.log
for d in 1 2 3; do for w in 1 2; do echo "== DBG=$d W=$w"; B200DP_RING=3 B200DP_DBG=$d B200DP_WARPS=$w timeout 100 python -u scripts/gpu_time.py 2>&1 | grep "^W=0"; done; done | tee gpurun_out/dbg_modes.log
echo "== pytest parity"; timeout 900 python -u -m pytest tests/test_gpu_parity.py -m gpu -q -x --maxfail=3 > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/pytest_parity.log
echo "== pytest api"; timeout 300 python -u -m pytest tests/test_gpu_api.py -m gpu -q --maxfail=5 > gpurun_out/pytest_api.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/pytest_api.log
echo "== smoke" ; timeout 120 python -u __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 300 python -u bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log | cut -c1-2500
if [ -n "$PROF" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?"; grep -c softdp gpurun_out/launches.csv
echo "== ncu full (fwd, bwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softdp_ -s 6 -c 2 -f -o gpurun_out/prof_fwd_bwd \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
