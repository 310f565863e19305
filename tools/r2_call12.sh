#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dp_sync_check.py --graph 1 > gpurun_out/c12_sync.log 2> gpurun_out/c12_sync.err
echo "graph sync rc=$?"; grep '^{' gpurun_out/c12_sync.log; grep -v "^\[W\|^W0\|^$" gpurun_out/c12_sync.err | grep -E "Error|error|File|line" | head -30
