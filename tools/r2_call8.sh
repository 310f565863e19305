#!/bin/bash
# 512-thread capsule agreement kernels; pair kernel with 256-wide N tiles
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "capsule or heads or route" > gpurun_out/c8_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c8_tests.log | tail -5
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -k "splitz_cap_ca or onlycapsule" > gpurun_out/c8_tests2.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c8_tests2.log | tail -5
for c in splitz_cap_ca onlycapsule; do
timeout 150 python bench.py --config $c --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$c', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
for bn in 128 256; do
  EKL_TC2=2 EKL_TC_BN=$bn timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c8_layers_pair_bn$bn.json > gpurun_out/c8_layers_pair_bn$bn.log 2>&1
  echo "layer_bench pair BN=$bn rc=$?"
done
EKL_TC2=0 EKL_TC_BN=256 timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c8_layers_single_bn256.json > gpurun_out/c8_layers_single_bn256.log 2>&1
EKL_TC2=0 EKL_TC_BN=128 timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c8_layers_single_bn128.json > gpurun_out/c8_layers_single_bn128.log 2>&1
EKL_TC2=0 EKL_TC_BN=64 timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c8_layers_single_bn64.json > gpurun_out/c8_layers_single_bn64.log 2>&1
echo done
