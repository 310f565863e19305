#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider -k "not multigpu" > gpurun_out/c4_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(PASS|FAIL)" gpurun_out/c4_tests.log | grep -c PASS
grep -E "^FAIL" gpurun_out/c4_tests.log | head -20
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c4_tests.log | tail -12
grep -E "eval-mode|D_NET.*input gradient" gpurun_out/c4_tests.log
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/c4_bench.log 2> gpurun_out/c4_bench.err; echo "bench rc=$?" )
tail -2 gpurun_out/c4_bench.err
grep '^{' gpurun_out/c4_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline'] or {}
print('value', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value']), '; roofline', r.get('kernel'), round(r.get('frac', 0), 3), 'launches/step', d['gpu_launches'] / d['steps'])
for k, v in (r.get('families') or {}).items(): print('   ', k, v['us_per_step'], v['launches_per_step'], v.get('tflops_reference_count'), v.get('gbs_algorithmic'))"
for cfg in splitz_cap_ca coco catcls onlycapsule; do
timeout 150 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$cfg', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
