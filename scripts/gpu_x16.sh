#!/bin/bash
# the chained forward under the diagnostic builds of gpu_x16.py (built here with
# B200DP_NVCC_EXTRA="-DB200DP_FWD3_DBG=d" python -m deepblast_b200.build, copied to scripts/_bin/)
cd "$(dirname "$0")/.."
python scripts/gpu_x16.py
for d in 8 9 10 1 2; do LIB=scripts/_bin/libb200dp_dbg$d.so python scripts/gpu_x16.py; done
