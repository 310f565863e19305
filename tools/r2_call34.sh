#!/bin/bash
# full GPU suite after the BatchNorm ring kernels and the sub-pixel resident-filter routing
set -u
mkdir -p gpurun_out
( timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=8 > gpurun_out/c34_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed|s call|s setup" gpurun_out/c34_tests.log | tail -14
