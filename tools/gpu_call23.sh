#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest23.log | cut -c1-300
for c in c3 c4 c5; do
  timeout 400 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench23_$c.json 2>gpurun_out/r2_bench23_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench23_$c.json').read().strip().splitlines()[-1]);print('$c',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()}, 'frac', round(d['roofline']['frac'],3), round(d['roofline']['k1_standalone_frac'],3))" || tail -5 gpurun_out/r2_bench23_$c.err
done
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench23_c2.json 2>gpurun_out/r2_bench23_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench23_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), round(d['ms_per_step']*1e3,1), {k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})
for k,v in d['ops'].items(): print('ops', k, round(v['ms']*1e3,1), 'us', round(v['frac'],3))
print('collective', d['collective']['value'], d['collective']['ms_per_step'])
print('cpu', d['cpu_baseline']['value'], 'e2e', d['e2e']['value'])
PY
python tools/prof_step.py c2 4 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"warp_decode_kernel|render_mse_kernel" -s 4 -c 2 -o gpurun_out/r2b_k1_k3_full python tools/prof_step.py c2 4 > gpurun_out/r2_prof_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2_prof_ncu.log
