#!/bin/bash
# One GPU-box visit: staged bring-up, parity tests, smoke, a short bench.  Every stage has
# its own `timeout` so a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== staged"; timeout 600 python -u scripts/gpu_stage.py v2 2>&1 | tee gpurun_out/stage.log
echo "== timing"; timeout 300 python -u scripts/gpu_time.py 2>&1 | tee gpurun_out/time_sweep.log
echo "== pytest parity"; timeout 900 python -u -m pytest tests/test_gpu_parity.py -m gpu -q -x --maxfail=3 > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/pytest_parity.log
echo "== pytest api"; timeout 300 python -u -m pytest tests/test_gpu_api.py -m gpu -q --maxfail=5 > gpurun_out/pytest_api.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/pytest_api.log
echo "== smoke" ; timeout 120 python -u __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 300 python -u bench.py --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log | cut -c1-1500
