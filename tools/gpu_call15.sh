#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_reference_step.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest15.log | cut -c1-400
for R in 2 8; do timeout 60 python tools/select_phases_multi.py $R 4352 2>&1 | grep -v Warning; done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench15_c2.json 2>gpurun_out/r2_bench15_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench15_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), round(d['ms_per_step']*1e3,1))
for k,v in d['ops'].items(): print('ops', k, round(v['ms']*1e3,1), 'us', round(v['frac'],3))
PY
