#!/bin/bash
# CTA-pair conv kernel: forced on every eligible kernel_check shape, then the whole-step parity cases, then A/B bench
set -u
mkdir -p gpurun_out
for g in tc_fwd tc_dgrad tc_split fold; do
  EKL_TC2=2 timeout 200 python tools/kernel_check.py --group $g > gpurun_out/c5_kc_$g.log 2>&1
  echo "group $g (forced pairs): $(grep -c '^PASS' gpurun_out/c5_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c5_kc_$g.log) fail; $(tail -1 gpurun_out/c5_kc_$g.log)"
  grep -E "^FAIL|timeout|error" gpurun_out/c5_kc_$g.log | head -8
done
( timeout 600 python -m pytest tests/test_step_parity_gpu.py tests/test_zz_generation_gpu.py -m gpu -q -s -p no:cacheprovider -k "3stages-4 or 3stages-24 or splitz_cap_ca-32 or two_head or onlycapsule-4 or coco-64 or graphed" > gpurun_out/c5_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c5_tests.log | tail -8
grep -E "img[0-9] |errG|G grad median|D2 grad median|G floor|D2 floor" gpurun_out/c5_tests.log | head -40
for m in 0 1; do
EKL_TC2=$m timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC2=$m 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
for m in 0 1; do
EKL_TC2=$m timeout 150 python bench.py --config coco --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC2=$m coco', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 150 python tools/step_profile.py --config 3stages --json gpurun_out/c5_prof_3stages.json > gpurun_out/c5_prof_3stages.log 2>&1; head -24 gpurun_out/c5_prof_3stages.log | cut -c1-150
