#!/bin/bash
set -u
mkdir -p gpurun_out
for r in 0 45 60 80 120 20; do
EKL_TC_RHO=$r timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC_RHO=$r 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
for c in coco splitz_cap_ca catcls; do
for r in 0 45 80; do
EKL_TC_RHO=$r timeout 150 python bench.py --config $c --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC_RHO=$r $c', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
