#!/bin/bash
# split-K threshold 2 as the default: parity at the BASELINE batches, split-K kernel group, coco A/B
set -u
mkdir -p gpurun_out
timeout 300 python tools/kernel_check.py --group tc_split > gpurun_out/c48_kc.log 2>&1
echo "group tc_split: $(grep -c '^PASS' gpurun_out/c48_kc.log) pass, $(grep -c '^FAIL' gpurun_out/c48_kc.log) fail"
( timeout 600 python -m pytest tests/test_step_parity_gpu.py tests/test_zz_generation_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/c48_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c48_tests.log | tail -5
for m in 3 2; do
EKL_TC_SPLIT_MIN=$m timeout 150 python bench.py --config coco --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC_SPLIT_MIN=$m coco', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
