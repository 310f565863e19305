#!/bin/bash
# capsule routing kernels, Adam overlapped with the discriminator backward, new pair-kernel rule
set -u
mkdir -p gpurun_out
for g in route misc; do
  timeout 300 python tools/kernel_check.py --group $g > gpurun_out/c7_kc_$g.log 2>&1
  echo "group $g: $(grep -c '^PASS' gpurun_out/c7_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c7_kc_$g.log) fail; $(tail -1 gpurun_out/c7_kc_$g.log)"
  grep -E "^FAIL|timeout|rror" gpurun_out/c7_kc_$g.log | head -8
done
( timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_zz_generation_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/c7_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c7_tests.log | tail -8
for t in 0 1; do
EKL_TAIL_ADAM=$t timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TAIL_ADAM=$t 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value']))"
done
EKL_TC2=0 timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC2=0 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
for c in splitz_cap_ca onlycapsule; do
timeout 150 python bench.py --config $c --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$c', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 150 python tools/step_profile.py --config splitz_cap_ca --json gpurun_out/c7_prof_cfg4.json > gpurun_out/c7_prof_cfg4.log 2>&1; grep -A22 "^step" gpurun_out/c7_prof_cfg4.log | cut -c1-150
