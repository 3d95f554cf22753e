#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 tools/k1_micro > gpurun_out/r2_k1_micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/r2_k1_micro.log
timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
for c in c4 c3; do for pf in 0 32 64; do
  UBPL_K1_PF_MB=$pf UBPL_BENCH_PREFETCH=$pf timeout 300 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2_${c}_pf$pf.json 2>gpurun_out/r2_bench2_${c}_pf$pf.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench2_${c}_pf$pf.json').read().strip().splitlines()[-1]);print('$c pf=$pf',d['value'],d['ms_per_step'],d['roofline']['stages_ms'])"
done; done
