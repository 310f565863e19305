#!/bin/bash
# ncu --set full of the BatchNorm kernels at the generator's big backward shapes (GLU): what limits them?
set -u
mkdir -p gpurun_out
timeout 120 python tools/bn_bench.py 5 393216,64,1,1 98304,128,1,1 1572864,32,1,1 294912,128,3,2 > gpurun_out/c22_bn_plain.log 2>&1
echo "plain rc=$?"; grep -v Warn gpurun_out/c22_bn_plain.log | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bn_act_(fwd|bwd_reduce|bwd_apply)_kernel' --launch-skip 9 -c 6 -f -o gpurun_out/c22_bn_full \
    python tools/bn_bench.py 1 393216,64,1,1 98304,128,1,1 > gpurun_out/c22_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/c22_bn_full.ncu-rep
