#!/bin/bash
# generator tail update (marks at the condition code / stem), K-split VC_NET backward
set -u
mkdir -p gpurun_out
timeout 300 python tools/kernel_check.py --group vc > gpurun_out/c17_kc_vc.log 2>&1
echo "group vc: $(grep -c '^PASS' gpurun_out/c17_kc_vc.log) pass, $(grep -c '^FAIL' gpurun_out/c17_kc_vc.log) fail"
( timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_zz_generation_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/c17_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c17_tests.log | tail -5
for rep in 1 2 3; do
for t in 0 1; do
EKL_TAIL_ADAM=$t timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_TAIL_ADAM=$t 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
