#!/bin/bash
# round-2 call 1: parity of the new K1 (cooperative exhaustive decode, fast reductions) + first timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -15 gpurun_out/r2_pytest1.log
for pf in 0 8 2 32; do
  UBPL_BENCH_PREFETCH=$pf UBPL_K1_PF_EVERY=$pf timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench1_pf$pf.json 2> gpurun_out/r2_bench1_pf$pf.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench1_pf$pf.json").read().strip().splitlines()[-1])
    print("pf=$pf", d["value"], d["ms_per_step"], d["roofline"]["stages_ms"], d["config"]["exhaustive_decode_frac"])
except Exception as e:
    print("pf=$pf failed", e); print(open("gpurun_out/r2_bench1_pf$pf.err").read()[-2000:])
PY
done
UBPL_BENCH_OVERLAP_EMA=k3 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench1_emak3.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench1_emak3.json').read().strip().splitlines()[-1]);print('ema@k3',d['value'],d['ms_per_step'],d['roofline']['stages_ms'])"
for c in c3 c4 c5; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench1_$c.json 2>gpurun_out/r2_bench1_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench1_$c.json').read().strip().splitlines()[-1]);print('$c',d['value'],d['ms_per_step'],d['roofline']['stages_ms'])"
done
