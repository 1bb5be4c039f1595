#!/bin/bash
mkdir -p gpurun_out
{
echo "== B=2048 (W=1 -> 14 warps/SM)"; timeout 200 python -u scripts/gpu_time.py 2048 256 256 2>&1 | grep -E "^W=|alternating"
echo "== B=4096"; timeout 200 python -u scripts/gpu_time.py 4096 256 256 2>&1 | grep -E "^W=0|^W=1 grid=0|alternating"
echo "== B=1024 512x512"; timeout 200 python -u scripts/gpu_time.py 1024 512 512 2>&1 | grep -E "^W=0|^W=1 grid=0|^W=2 grid=0|alternating"
} 2>&1 | tee gpurun_out/dbg.log
