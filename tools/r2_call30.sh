#!/bin/bash
# sub-pixel plans on the resident-filter kernel: correctness, per-layer timing A/B, step A/B
set -u
mkdir -p gpurun_out
for g in tc_fwd tc_dgrad; do
timeout 400 python tools/kernel_check.py --group $g > gpurun_out/c30_kc_$g.log 2>&1
echo "group $g: $(grep -c '^PASS' gpurun_out/c30_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c30_kc_$g.log) fail"; grep '^FAIL' gpurun_out/c30_kc_$g.log | head -12
done
for rep in 1 2; do
for e in 0 1; do
EKL_RW_SUBPIXEL=$e timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_RW_SUBPIXEL=$e 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or catcls-4" > gpurun_out/c30_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c30_tests.log | tail -5
