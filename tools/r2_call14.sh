#!/bin/bash
set -u
for rep in 1 2 3; do
for r in 0 5 1 4; do
EKL_BN_REV=$r timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_BN_REV=$r 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
