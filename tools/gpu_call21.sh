#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # label, env...
  lab=$1; shift
  env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_b21_$lab.json 2>/dev/null
  python -c "
import json;d=json.loads(open('gpurun_out/r2_b21_$lab.json').read().strip().splitlines()[-1]);print('$lab',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
}
run base A=1
run ema4 UBPL_EMA_CTAS=4
run ema2 UBPL_EMA_CTAS=2
run ema1 UBPL_EMA_CTAS=1
run cap6 UBPL_K1_INFLIGHT=6
run cap6ema2 UBPL_K1_INFLIGHT=6 UBPL_EMA_CTAS=2
run cap12 UBPL_K1_INFLIGHT=12
run nopf UBPL_BENCH_PREFETCH=0
run pf96 UBPL_K1_PF_MB=96
run pfe4 UBPL_K1_PF_EVERY=4
run pfe16 UBPL_K1_PF_EVERY=16
