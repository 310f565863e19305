#!/bin/bash
set -u
mkdir -p gpurun_out
EKL_BN_REV=7 timeout 200 python tools/kernel_check.py --group bn > gpurun_out/c13_kc_bn_rev7.log 2>&1
echo "group bn (rev=7): $(grep -c '^PASS' gpurun_out/c13_kc_bn_rev7.log) pass, $(grep -c '^FAIL' gpurun_out/c13_kc_bn_rev7.log) fail"
for r in 0 1 4 5 2 3; do
EKL_BN_REV=$r timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_BN_REV=$r 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
for r in 0 5 3; do
EKL_BN_REV=$r timeout 150 python bench.py --config coco --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_BN_REV=$r coco', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
for r in 0 5; do
  EKL_BN_REV=$r timeout 300 python tools/layer_bench.py --config 3stages --only bn --iters 7 --json gpurun_out/c13_layers_bn_rev$r.json > gpurun_out/c13_layers_bn_rev$r.log 2>&1
done
