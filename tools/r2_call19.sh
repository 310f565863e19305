#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or coco-64 or catcls-4" > gpurun_out/c19_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c19_tests.log | tail -5
for rep in 1 2 3; do
for t in 0 1; do
EKL_FC_T=$t timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_FC_T=$t 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
for t in 0 1; do
EKL_FC_T=$t timeout 150 python bench.py --config coco --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_FC_T=$t coco', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 150 python tools/step_profile.py --config 3stages --json gpurun_out/c19_prof_3stages.json > gpurun_out/c19_prof_3stages.log 2>&1; grep -E "cutlass|gemm|gemv" gpurun_out/c19_prof_3stages.log | cut -c1-170
