#!/bin/bash
# the training-shaped step of bench.py alone (CUDA graph time), N times
cd "$(dirname "$0")/.."
for i in 1 2; do
python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
w = d['workloads']
print('C2 %.1f G' % (d['value'] / 1e9), 'train_step_c2', w['train_step_c2'].get('graph_ms'), w['train_step_c2'].get('eager_ms'), 'b32x512', w['train_step_b32x512'].get('graph_ms'))
"
done
