#!/bin/bash
# narrow-strip small BatchNorm backward; LeakyReLU' of the discriminator stem fused into the next conv's data-gradient
set -u
mkdir -p gpurun_out
for g in bn tc_dgrad; do
timeout 400 python tools/kernel_check.py --group $g > gpurun_out/c41_kc_$g.log 2>&1
echo "group $g: $(grep -c '^PASS' gpurun_out/c41_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c41_kc_$g.log) fail"; grep '^FAIL' gpurun_out/c41_kc_$g.log | head -12
done
timeout 100 python tools/bn_bench.py 5 1152,512,3,2 1152,1024,3,2 384,512,1,2 4608,1024,3,2 2>&1 | grep -v Warn | tail -4 | cut -c1-60,120-200
for rep in 1 2 3; do
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or splitz_cap_ca-32" > gpurun_out/c41_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c41_tests.log | tail -5
