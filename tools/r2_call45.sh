#!/bin/bash
# final validation of the round: full GPU suite, smoke(), default bench line (all configs), CUPTI step profile
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/c45_tests.log 2>&1; echo "pytest rc=$?" )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c45_tests.log | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/c45_bench.log 2> gpurun_out/c45_bench.err
echo "bench rc=$?"; grep '^{' gpurun_out/c45_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'roof', round(r['frac'], 3), 'launches', d['gpu_launches'])
print({k: round(v['value']) for k, v in d['extra']['all_configs'].items()})"
timeout 200 python tools/step_profile.py --config 3stages --json gpurun_out/r02_step_profile_3stages_final.json > gpurun_out/r02_step_profile_3stages_final.log 2>&1
echo "step_profile rc=$?"; grep -v Warn gpurun_out/r02_step_profile_3stages_final.log | sed -n 3,4p | cut -c1-120
