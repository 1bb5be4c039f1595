#!/bin/bash
mkdir -p gpurun_out
{
echo "== single process"; timeout 120 python -u scripts/dbg_2gpu.py
echo "== single process OMP=1"; OMP_NUM_THREADS=1 timeout 120 python -u scripts/dbg_2gpu.py
echo "== torchrun 2, with PG"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dbg_2gpu.py
echo "== torchrun 2, no PG"; NO_PG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/dbg_2gpu.py
} 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS env\|^$" | tee gpurun_out/dbg2.log
