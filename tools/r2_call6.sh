#!/bin/bash
# fold / vc / bn kernel groups, the device image pyramid, per-layer A/B of the CTA-pair conv kernel and of BN rows in flight
set -u
mkdir -p gpurun_out
for g in fold vc bn misc; do
  timeout 200 python tools/kernel_check.py --group $g > gpurun_out/c6_kc_$g.log 2>&1
  echo "group $g: $(grep -c '^PASS' gpurun_out/c6_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c6_kc_$g.log) fail; $(tail -1 gpurun_out/c6_kc_$g.log)"
  grep -E "^FAIL|timeout|rror" gpurun_out/c6_kc_$g.log | head -8
done
EKL_BN_U=4 timeout 200 python tools/kernel_check.py --group bn > gpurun_out/c6_kc_bn_u4.log 2>&1
echo "group bn (U=4): $(grep -c '^PASS' gpurun_out/c6_kc_bn_u4.log) pass, $(grep -c '^FAIL' gpurun_out/c6_kc_bn_u4.log) fail"
( timeout 300 python -m pytest tests/test_zz_generation_gpu.py -m gpu -q -p no:cacheprovider -k "pyramid or two_head" > gpurun_out/c6_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c6_tests.log | tail -8
for m in 0 2; do
  EKL_TC2=$m timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c6_layers_tc2_$m.json > gpurun_out/c6_layers_tc2_$m.log 2>&1
  echo "layer_bench EKL_TC2=$m rc=$?"; tail -3 gpurun_out/c6_layers_tc2_$m.log | cut -c1-160
done
for u in 2 4; do
  EKL_BN_U=$u timeout 300 python tools/layer_bench.py --config 3stages --only bn --iters 7 --json gpurun_out/c6_layers_bn_u$u.json > gpurun_out/c6_layers_bn_u$u.log 2>&1
  echo "layer_bench EKL_BN_U=$u rc=$?"; tail -3 gpurun_out/c6_layers_bn_u$u.log | cut -c1-160
done
for u in 2 4; do
EKL_BN_U=$u timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_BN_U=$u 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value']))"
done
