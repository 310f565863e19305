#!/bin/bash
# leaner BatchNorm streaming kernels: correctness (kernel group bn + two whole-step parity cases), timing, instruction counts
set -u
mkdir -p gpurun_out
timeout 300 python tools/kernel_check.py --group bn > gpurun_out/c24_kc_bn.log 2>&1
echo "group bn: $(grep -c '^PASS' gpurun_out/c24_kc_bn.log) pass, $(grep -c '^FAIL' gpurun_out/c24_kc_bn.log) fail"; grep '^FAIL' gpurun_out/c24_kc_bn.log | head
timeout 120 python tools/bn_bench.py 5 393216,64,1,1 98304,128,1,1 1572864,32,1,1 294912,128,3,2 393216,32,1,0 > gpurun_out/c24_bn_plain.log 2>&1
grep -v Warn gpurun_out/c24_bn_plain.log | tail -5
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or splitz_cap_ca-4" > gpurun_out/c24_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c24_tests.log | tail -5
for rep in 1 2; do
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'bn_act_(fwd|bwd_reduce|bwd_apply)_kernel' --launch-skip 9 -c 6 -f -o gpurun_out/c24_bn_full \
    python tools/bn_bench.py 1 393216,64,1,1 98304,128,1,1 > gpurun_out/c24_ncu.log 2>&1
echo "ncu rc=$?"
