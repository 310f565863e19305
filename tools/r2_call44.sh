#!/bin/bash
# fused split-K finish + BatchNorm forward; VC_NET forward with 512-column chunks
set -u
mkdir -p gpurun_out
for g in tc_split vc; do
timeout 400 python tools/kernel_check.py --group $g > gpurun_out/c44_kc_$g.log 2>&1
echo "group $g: $(grep -c '^PASS' gpurun_out/c44_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c44_kc_$g.log) fail"; grep '^FAIL' gpurun_out/c44_kc_$g.log | head -12 | cut -c1-400
done
grep "fused-bn" gpurun_out/c44_kc_tc_split.log | head -2 | cut -c1-300
for rep in 1 2; do
for e in 0 1; do
EKL_SPLIT_BN=$e timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_SPLIT_BN=$e 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or splitz_cap_ca-32 or catcls-4" > gpurun_out/c44_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c44_tests.log | tail -5
