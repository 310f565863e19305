#!/bin/bash
# margins of the floor-relative parity bounds at the BASELINE batches (two runs: the step has fp32 atomics)
set -u
mkdir -p gpurun_out
for rep in 1 2; do
timeout 700 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -p no:cacheprovider -s -k "3stages-24 or catcls-24 or onlycapsule-32 or splitz_cap_ca-32 or coco-64" > gpurun_out/c43_parity_$rep.log 2>&1
echo "rep $rep rc=$?"; grep -E "passed|failed" gpurun_out/c43_parity_$rep.log | tail -2
done
