#!/bin/bash
# split operand refresh (forward packs first): conv kernel groups, whole-step parity subset, bench, step trace
set -u
mkdir -p gpurun_out
for g in tc_fwd tc_dgrad fold; do
  timeout 300 python tools/kernel_check.py --group $g > gpurun_out/c16_kc_$g.log 2>&1
  echo "group $g: $(grep -c '^PASS' gpurun_out/c16_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c16_kc_$g.log) fail"
done
( timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_zz_generation_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or splitz_cap_ca-32 or coco-64 or graphed or eval or checkpoint or variants" > gpurun_out/c16_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c16_tests.log | tail -5
for rep in 1 2; do
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']))"
done
timeout 200 python tools/graph_modes.py --captures 2 2>/dev/null | grep capture
timeout 200 python tools/step_trace.py --config 3stages --json gpurun_out/c16_trace.json > gpurun_out/c16_trace.log 2>&1
grep -A12 "^step span" gpurun_out/c16_trace.log | cut -c1-150
