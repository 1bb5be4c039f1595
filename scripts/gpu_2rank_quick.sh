#!/bin/bash
# 2-rank torchrun smoke of bench.py at the final commit (needs gpurun --gpus 2): c2 and the reference arm.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_c2.log 2>&1; echo "c2 rc=$?"
tail -1 gpurun_out/bench_2gpu_c2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('n_gpus %d value %.1f G  ms/step %.3f  fwd %.3f bwd %.3f  e2e %.2f G (%s)' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], r['fwd']['ms'], r['bwd']['ms'], d['e2e']['value']/1e9, d['e2e']['host_affinity']))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 3 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "ref rc=$?"
tail -1 gpurun_out/bench_2gpu_ref.log | cut -c1-160
