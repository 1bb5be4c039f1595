#!/bin/bash
mkdir -p gpurun_out
echo "== staged v3"; timeout 900 python -u scripts/gpu_stage.py v3 2>&1 | tee gpurun_out/stage_v3.log
for w in "$@"; do
echo "== timing v3 $w"; timeout 600 python -u scripts/gpu_time_v3.py $w 2>&1 | tee gpurun_out/time_v3_$w.log
done
