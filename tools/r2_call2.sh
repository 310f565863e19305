#!/bin/bash
# round-2 second GPU call (2 GPUs): N=1 vs N=2 bench under the driver's command, bf16 vs fp32 gradient payload, timeline
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
run2() { # $1 tag, rest: env assignments
  tag=$1; shift
  ( env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c2_$tag.log 2> gpurun_out/c2_$tag.err; echo "$tag rc=$?" )
  grep '^{' gpurun_out/c2_$tag.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline'] or {}
print('$tag', 'n', d['n_gpus'], round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value']), d['config'].get('grad_comm'), 'roof', round(r.get('frac', 0), 3))
for k, v in (r.get('families') or {}).items(): print('   ', k, v['us_per_step'], v['launches_per_step'])"
  tail -2 gpurun_out/c2_$tag.err
}
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/c2_n1.log 2> gpurun_out/c2_n1.err; echo "n1 rc=$?" )
grep '^{' gpurun_out/c2_n1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('n1', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value']))"
run2 bf16 EKL_GRAD_COMM=bf16
run2 fp32 EKL_GRAD_COMM=fp32

( timeout 300 python -m pytest tests/test_zz_multigpu.py -m gpu -q -x 2>&1 | tail -3 )
timeout 120 python tools/step_trace.py --config 3stages --json gpurun_out/c2_trace_3stages.json > gpurun_out/c2_trace_3stages.log 2>&1; tail -60 gpurun_out/c2_trace_3stages.log
