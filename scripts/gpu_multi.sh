#!/bin/bash
# Workload sweep on one GPU + a 2-rank torchrun run (needs gpurun --gpus 2).
mkdir -p gpurun_out
echo "== timing"; timeout 200 python -u scripts/gpu_time.py 2>&1 | grep -E "^W=0|^W=1 grid=0|alternating" | tee gpurun_out/time_sweep.log
for w in c2 c3 c4 c5; do
  echo "== bench $w"; timeout 300 python -u bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_$w.log 2>&1; echo "rc=$?"
  tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('value %.1f G  ms/step %.3f  fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)  e2e %.2f G  cpu %.3f G x%d cores' % (d['value']/1e9, d['ms_per_step'], r['fwd']['ms'], r['fwd']['frac'], r['bwd']['ms'], r['bwd']['frac'], d['e2e']['value']/1e9, d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores']))
except Exception as e: print('parse fail', e)
"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== torchrun 2 ranks"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?"
  tail -2 gpurun_out/bench_2gpu.log | cut -c1-600
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 2 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "ref rc=$?"
  tail -1 gpurun_out/bench_2gpu_ref.log | cut -c1-300
fi
