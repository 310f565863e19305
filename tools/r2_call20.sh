#!/bin/bash
# end-of-round evidence: default bench line, launch list, --set full of the top kernels, CUPTI step profile
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/c20_bench.log 2> gpurun_out/c20_bench.err
echo "bench rc=$?"; grep '^{' gpurun_out/c20_bench.log | cut -c1-1500
timeout 200 python tools/step_profile.py --config 3stages --json gpurun_out/r02_step_profile_3stages_final.json > gpurun_out/r02_step_profile_3stages_final.log 2>&1
echo "step_profile rc=$?"; head -30 gpurun_out/r02_step_profile_3stages_final.log | cut -c1-160
bash tools/ncu_refresh.sh r02 list
bash tools/ncu_refresh.sh r02 full
