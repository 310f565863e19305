#!/bin/bash
# 2 GPUs: the multi-GPU tests (bench under the driver's command, replicas bit-identical) and the N = 2 bench line
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_zz_multigpu.py -m gpu -q -p no:cacheprovider > gpurun_out/c11_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed|skipped" gpurun_out/c11_tests.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/c11_n2.log 2> gpurun_out/c11_n2.err
echo "N=2 rc=$?"; grep '^{' gpurun_out/c11_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=2', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']), 'roof', d['roofline'] and round(d['roofline']['frac'], 3))"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/c11_n1.log 2> gpurun_out/c11_n1.err
echo "N=1 rc=$?"; grep '^{' gpurun_out/c11_n1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=1', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']), 'roof', d['roofline'] and round(d['roofline']['frac'], 3))"
