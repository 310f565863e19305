#!/bin/bash
set -u
mkdir -p gpurun_out
for g in ; do
timeout 400 python tools/kernel_check.py --group $g > gpurun_out/c35_kc_$g.log 2>&1
echo "group $g: $(grep -c '^PASS' gpurun_out/c35_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c35_kc_$g.log) fail"; grep '^FAIL' gpurun_out/c35_kc_$g.log | head -12
done
for shape in "0 24 64 64 64 128 dgrad"; do
  for e in 1 0; do
    echo -n "DISABLE_RW=$e  "; EKL_DISABLE_RW=$e timeout 60 python tools/conv_one.py $shape 5 2>&1 | tail -1
  done
done
for rep in 1 2 3; do
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24" > gpurun_out/c35_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c35_tests.log | tail -5
