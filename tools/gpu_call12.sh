#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "multirank" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest11.log | cut -c1-300
timeout 200 python tools/k3_variants.py > gpurun_out/r2_k3_var.log 2>&1; echo "k3 rc=$?"; grep -v Warning gpurun_out/r2_k3_var.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench12_c2.json 2>gpurun_out/r2_bench12_c2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench12_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench12_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), round(d['ms_per_step']*1e3,1), {k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})
print('twin', d['roofline']['stages_note'][:200])
print('sens', json.dumps(d.get('sensitivity'), indent=0))
print('ops', json.dumps(d.get('ops'), indent=0))
print('cpu', d.get('cpu_baseline')); print('e2e', d['e2e'])
PY
