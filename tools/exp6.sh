cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/s2_pytest4.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/s2_pytest4.log
UBPL_AB_MASKS=0 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 c4 c3 > gpurun_out/s2_k1_v4.log 2>&1
cat gpurun_out/s2_k1_v4.log
for ov in k1 tail; do
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras > gpurun_out/s2_b6_$ov.json 2> gpurun_out/s2_b6_$ov.err
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras --config c3 > gpurun_out/s2_b6_c3_$ov.json 2> gpurun_out/s2_b6_c3_$ov.err
done
UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras --config c4 > gpurun_out/s2_b6_c4_tail.json 2> gpurun_out/s2_b6_c4_tail.err
UBPL_K1_PF_MB=0 UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras > gpurun_out/s2_b6_tail_nopf.json 2> gpurun_out/s2_b6_tail_nopf.err
for f in gpurun_out/s2_b6_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step']*1e3,1), {k:round(v*1e3,1) for k,v in d['roofline']['stages_ms'].items() if v is not None})
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
