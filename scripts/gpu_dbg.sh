#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
  for w in 1 2; do
    echo "== DBG=$d W=$w"; B200DP_DBG=$d B200DP_WARPS=$w timeout 120 python -u scripts/gpu_time.py 2>&1 | grep "W=0"
  done
done 2>&1 | tee gpurun_out/dbg.log
