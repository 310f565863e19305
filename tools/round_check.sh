#!/bin/bash
# One-call GPU validation of a round's state, meant for `gpurun --timeout 600 -- 'bash tools/round_check.sh'`:
# the GPU test suite, smoke(), the default bench line, one A/B of the scheduling knobs and the per-kernel step profile.
# Everything is wrapped in `timeout`; outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/rc_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee gpurun_out/rc_smoke.log
timeout 200 python bench.py > gpurun_out/rc_bench.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/rc_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value']), 'img/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value']),
      '; roofline frac', round(d['roofline']['frac'], 3), '; cpu', d.get('cpu_baseline', {}).get('value'),
      '; gpu eager', {k: round(v) for k, v in d.get('gpu_eager_baseline', {}).items() if isinstance(v, float)})"
EKL_PARALLEL_D=0 timeout 90 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('serial discriminators:', round(d['value']), 'img/s')"
EKL_D_PRIO=1 timeout 90 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('high-priority deepest branch (EKL_D_PRIO=1):', round(d['value']), 'img/s')"
EKL_WGRAD_STREAM=1 timeout 90 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('weight gradients on side streams (EKL_WGRAD_STREAM=1):', round(d['value']), 'img/s')"
EKL_WGRAD_STREAM=1 timeout 120 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -k "3stages or splitz" 2>&1 | tail -2
# 2-GPU experiments (run under gpurun --gpus 2, one at a time, always inside `timeout`):
#   [EKL_BUCKET_AR=1] timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-profile
timeout 120 python tools/step_profile.py --config 3stages > gpurun_out/rc_step_profile.log 2>&1; tail -25 gpurun_out/rc_step_profile.log
