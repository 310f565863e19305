#!/bin/bash
# round-2 first GPU call: full GPU suite (with the per-tensor parity report), default bench, scheduling A/Bs, step profiles
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
( timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/c1_tests.log 2>&1; echo "pytest rc=$?" ) 
tail -15 gpurun_out/c1_tests.log
grep -E "passed|failed" gpurun_out/c1_tests.log | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/c1_smoke.log
( timeout 420 python bench.py > gpurun_out/c1_bench.log 2> gpurun_out/c1_bench.err; echo "bench rc=$?" )
tail -3 gpurun_out/c1_bench.err
grep '^{' gpurun_out/c1_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline'] or {}
print('value', round(d['value']), 'img/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value']), '; roofline', r.get('kernel'), round(r.get('frac', 0), 3), 'exec', round(r.get('achieved_executed', 0)))
print('cpu', d.get('cpu_baseline')); print('eager', d.get('gpu_eager_baseline')); print('extra', json.dumps(d.get('extra')))
for k, v in (r.get('families') or {}).items(): print('   ', k, v)
print('all conv', r.get('all_conv_kernels'))"
for kv in EKL_PARALLEL_D=0 EKL_D_PRIO=1 EKL_WGRAD_STREAM=1; do
  env $kv timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile --no-extra 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$kv:', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' | cut -c1-600
timeout 120 python tools/step_profile.py --config 3stages --json gpurun_out/c1_prof_3stages.json > gpurun_out/c1_prof_3stages.log 2>&1; head -45 gpurun_out/c1_prof_3stages.log
timeout 120 python tools/step_profile.py --config splitz_cap_ca --json gpurun_out/c1_prof_cfg4.json > gpurun_out/c1_prof_cfg4.log 2>&1; head -40 gpurun_out/c1_prof_cfg4.log
