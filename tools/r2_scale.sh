#!/bin/bash
# bench.py under the driver's torchrun command on N GPUs (N = $1)
set -u
N=${1:-4}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/scale_n$N.log 2> gpurun_out/scale_n$N.err
echo "N=$N rc=$?"; grep '^{' gpurun_out/scale_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=$N', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']), 'roof', d['roofline'] and round(d['roofline']['frac'], 3), d['clocks'])"
EKL_GRAD_COMM=fp32 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-extra --no-profile > gpurun_out/scale_n${N}_fp32.log 2> gpurun_out/scale_n${N}_fp32.err
echo "N=$N fp32 rc=$?"; grep '^{' gpurun_out/scale_n${N}_fp32.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=$N fp32 payload', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
