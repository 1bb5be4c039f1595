#!/bin/bash
# Full GPU validation: pytest -m gpu, smoke, bench c2 (+ ncu evidence when PROF=1).
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -u -m pytest tests -m gpu -q -x --maxfail=3 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 120 python -u __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
for w in ${WL:-c2}; do
echo "== bench $w"; timeout 300 python -u bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/bench_$w.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('value %.1f G  ms/step %.3f  fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)  e2e %.2f G  cpu %.3f G x%d cores  first3 %s max %.3f' % (d['value']/1e9, d['ms_per_step'], r['fwd']['ms'], r['fwd']['frac'], r['bwd']['ms'], r['bwd']['frac'], d['e2e']['value']/1e9, d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores'], d.get('step_ms_first3'), d.get('step_ms_max',0)))
except Exception as e: print('parse fail', e)
"
done
if [ -n "$PROF" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?"; grep -c softdp gpurun_out/launches.csv
echo "== ncu full (fwd, bwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softdp_ -s 6 -c 2 -f -o gpurun_out/prof_fwd_bwd \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
