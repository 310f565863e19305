#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for f in 0 1; do
for n in 0 1; do
EKL_LRELU_FUSE=$f EKL_BN_SMALL2=$n timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep LRELU_FUSE=$f BN_SMALL2=$n 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
done
