#!/bin/bash
# Refresh the committed ncu evidence, ONE profiler pass per GPU call:
#   gpurun --timeout 900 -- 'bash tools/ncu_refresh.sh r02 list'   launch list of one eager step -> gpurun_out/<tag>_launches_3stages.csv
#   gpurun --timeout 900 -- 'bash tools/ncu_refresh.sh r02 full'   --set full of the top kernels -> gpurun_out/<tag>_conv_full.ncu-rep
# each runs the plain command first (it must exit 0 before anything is profiled);
# then, back in the container (no GPU needed):
#   python tools/ncu_summary.py launches gpurun_out/<tag>_launches_3stages.csv "<cmd>" > profiles/<tag>_launches_3stages_summary.md
#   python tools/ncu_summary.py full gpurun_out/<tag>_conv_full.ncu-rep profiles/<tag>_conv_full.json > profiles/<tag>_conv_full_summary.md
# (bench.py reads profiles/r01_conv_full.json for roofline.traffic: point ncu_traffic() at the new file.)
# Single GPU only; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02}
MODE=${2:-list}
CMD="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-profile"
mkdir -p gpurun_out
timeout 120 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain command failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
if [ "$MODE" = list ]; then
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv \
    --log-file gpurun_out/${TAG}_launches_3stages.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${TAG}_launches_3stages.csv)"
else
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'conv_gemm_tc2?_kernel|conv3x3_rw_kernel|conv_wgrad_co_kernel|conv_wgrad_halo_kernel|bn_act_(fwd|bwd_reduce|bwd_apply)_kernel' \
    --launch-skip 40 -c 40 -f -o gpurun_out/${TAG}_conv_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full set rc=$?"; ls -la gpurun_out/${TAG}_conv_full.ncu-rep
fi
