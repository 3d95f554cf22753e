#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest13.log | cut -c1-300
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench13_c2.json 2>gpurun_out/r2_bench13_c2.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2_bench13_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench13_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), round(d['ms_per_step']*1e3,1), {k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})
for k,v in d['sensitivity'].items(): print('sens', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
