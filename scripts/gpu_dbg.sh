#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u scripts/gpu_timeline.py > gpurun_out/timeline.log 2>&1; echo rc=$?
head -120 gpurun_out/timeline.log
