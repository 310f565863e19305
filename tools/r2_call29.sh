#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2 3 4; do
for m in 3 2; do
EKL_TC_SPLIT_MIN=$m timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_TC_SPLIT_MIN=$m 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
