#!/bin/bash
set -u
mkdir -p gpurun_out
for g in 0 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$g tools/dp_sync_check.py --graph $g > gpurun_out/c18_sync$g.log 2> gpurun_out/c18_sync$g.err
echo "sync graph=$g rc=$?"; grep '^{' gpurun_out/c18_sync$g.log
done
EKL_GRAD_COMM=fp32 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/dp_sync_check.py --graph 1 --config splitz_cap_ca > gpurun_out/c18_sync2.log 2> gpurun_out/c18_sync2.err
echo "sync cfg4 fp32 rc=$?"; grep '^{' gpurun_out/c18_sync2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/c18_n2.log 2> gpurun_out/c18_n2.err
echo "N=2 rc=$?"; grep '^{' gpurun_out/c18_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('N=2', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']), 'roof', round(r['frac'], 3), 'match', r.get('launches_match_calls'), r.get('calls_accounted'), r.get('launches_per_step'))"
