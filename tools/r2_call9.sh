#!/bin/bash
# tile-width cost model with a total-traffic term: sweep of the fabric / per-SM ingest ratio
set -u
mkdir -p gpurun_out
for r in 0 25 35 45 60; do
  EKL_TC_RHO=$r timeout 300 python tools/layer_bench.py --config 3stages --only conv --iters 7 --json gpurun_out/c9_layers_rho$r.json > gpurun_out/c9_layers_rho$r.log 2>&1
  echo "layer_bench rho=$r rc=$?"
done
for r in 0 35 45; do
EKL_TC_RHO=$r timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EKL_TC_RHO=$r 3stages', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms')"
done
