#!/bin/bash
# kernels after the BN-statistics / code-fold / prepack / chained-branch changes: kernel groups, whole-step parity, bench, timeline
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider -k "not multigpu" > gpurun_out/c3_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(PASS|FAIL)" gpurun_out/c3_tests.log | grep -c PASS
grep -E "^FAIL|Error|error" gpurun_out/c3_tests.log | head -20
tail -5 gpurun_out/c3_tests.log
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/c3_bench.log 2> gpurun_out/c3_bench.err; echo "bench rc=$?" )
tail -3 gpurun_out/c3_bench.err
grep '^{' gpurun_out/c3_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline'] or {}
print('value', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value']), '; roofline', r.get('kernel'), round(r.get('frac', 0), 3), 'launches/step', d['gpu_launches'] / d['steps'])
for k, v in (r.get('families') or {}).items(): print('   ', k, v['us_per_step'], v['launches_per_step'], v.get('tflops_reference_count'), v.get('gbs_algorithmic'))"
for cfg in splitz_cap_ca coco; do
timeout 150 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$cfg', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 120 python tools/step_trace.py --config 3stages --json gpurun_out/c3_trace_3stages.json > gpurun_out/c3_trace_3stages.log 2>&1; grep -E "step span|stream |time with" gpurun_out/c3_trace_3stages.log
timeout 120 python tools/step_profile.py --config 3stages --json gpurun_out/c3_prof_3stages.json > gpurun_out/c3_prof_3stages.log 2>&1; head -30 gpurun_out/c3_prof_3stages.log | cut -c1-150
