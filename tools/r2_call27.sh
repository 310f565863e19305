#!/bin/bash
set -u
mkdir -p gpurun_out
for c in 0 1 2 3; do
echo "EKL_BN_RING=$c"
EKL_BN_RING=$c timeout 120 python tools/bn_bench.py 5 393216,64,1,1 98304,128,1,1 1572864,32,1,1 294912,128,3,2 393216,32,1,0 73728,256,3,2 18432,512,3,2 2>&1 | grep -v Warn | tail -7 | cut -c1-60,88-200
done
for rep in 1 2; do
for c in 0 1 2 3; do
EKL_BN_RING=$c timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_BN_RING=$c 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
