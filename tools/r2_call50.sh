#!/bin/bash
# traversal direction of the BatchNorm streaming passes with the ring kernels (bit 0 forward, 1 backward reduce, 2 backward apply)
set -u
for r in 0 4 0 4 5; do
EKL_BN_REV=$r timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_BN_REV=$r 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
