#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest9.log
UBPL_AB_CAPS=8,0 UBPL_AB_MASKS=0,4 timeout 150 python tools/k1_ab.py c2 c4 > gpurun_out/r2_k1_ab4.log 2>&1; echo "ab rc=$?"; grep -v Warning gpurun_out/r2_k1_ab4.log
timeout 300 python bench.py --config c2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench9_c2.json 2>gpurun_out/r2_bench9_c2.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench9_c2.json').read().strip().splitlines()[-1]);print('c2',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
