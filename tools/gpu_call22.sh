#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest22.log | cut -c1-300
run() { lab=$1; shift
  env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_b22_$lab.json 2>gpurun_out/r2_b22_$lab.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2_b22_$lab.json').read().strip().splitlines()[-1]);print('$lab',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})" || tail -5 gpurun_out/r2_b22_$lab.err
}
run late_pdl A=1
run late_nopdl UBPL_K3_PDL=0
run early_pdl UBPL_EMA_JOIN=early
run early_nopdl UBPL_EMA_JOIN=early UBPL_K3_PDL=0
