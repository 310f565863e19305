#!/bin/bash
# L2 chunks in the resident-filter kernel, class-head Linear as a batched split-K GEMM, parallel logit-head backward
set -u
mkdir -p gpurun_out
for g in tcchunk_fwd tcchunk_dgrad heads; do
timeout 400 python tools/kernel_check.py --group $g > gpurun_out/c38_kc_$g.log 2>&1
echo "group $g: $(grep -c '^PASS' gpurun_out/c38_kc_$g.log) pass, $(grep -c '^FAIL' gpurun_out/c38_kc_$g.log) fail"; grep '^FAIL' gpurun_out/c38_kc_$g.log | head -12
done
for shape in "2 72 128 128 64 128 dgrad"; do
  for e in 4000 40; do
    echo -n "CHUNK_MB=$e  "; EKL_RW_CHUNK_MB=$e timeout 60 python tools/conv_one.py $shape 5 2>&1 | tail -1
  done
done
for rep in 1 2; do
for f in 0 16; do
EKL_FC_SPLIT=$f timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rep $rep EKL_FC_SPLIT=$f 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
done
EKL_RW_CHUNK_MB=4000 timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('no chunks, FC_SPLIT=16 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
( timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "3stages-24 or coco-64" > gpurun_out/c38_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c38_tests.log | tail -5
