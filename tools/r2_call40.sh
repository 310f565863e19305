#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/kernel_check.py --group bn > gpurun_out/c40_kc_bn.log 2>&1
echo "group bn: $(grep -c '^PASS' gpurun_out/c40_kc_bn.log) pass, $(grep -c '^FAIL' gpurun_out/c40_kc_bn.log) fail"; grep '^FAIL' gpurun_out/c40_kc_bn.log | head
timeout 100 python tools/bn_bench.py 5 1152,512,3,2 1152,1024,3,2 384,512,1,2 4608,1024,3,2 2>&1 | grep -v Warn | tail -4 | cut -c1-60,120-200
for rep in 1 2 3; do
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
