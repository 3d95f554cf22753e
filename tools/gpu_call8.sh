#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in c2 c3 c4 c5; do
  timeout 300 python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench8_$c.json 2>gpurun_out/r2_bench8_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8_$c.json').read().strip().splitlines()[-1]);print('$c',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
done
UBPL_BENCH_OVERLAP_EMA=k3 timeout 300 python bench.py --config c2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench8_c2_emak3.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8_c2_emak3.json').read().strip().splitlines()[-1]);print('c2 ema@k3',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
UBPL_BENCH_OVERLAP_EMA=0 timeout 300 python bench.py --config c2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench8_c2_ema0.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8_c2_ema0.json').read().strip().splitlines()[-1]);print('c2 ema after',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
python tools/prof_step.py c2 4 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:warp_decode_kernel -s 2 -c 1 -o gpurun_out/r2a_k1_full python tools/prof_step.py c2 4 > gpurun_out/r2_prof_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_prof_ncu.log
