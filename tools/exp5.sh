cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest3.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/s2_pytest3.log
for ov in k1 tail; do
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras > gpurun_out/s2_b5_$ov.json 2> gpurun_out/s2_b5_$ov.err
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras --config c4 > gpurun_out/s2_b5_c4_$ov.json 2> gpurun_out/s2_b5_c4_$ov.err
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras --config c3 > gpurun_out/s2_b5_c3_$ov.json 2> gpurun_out/s2_b5_c3_$ov.err
done
UBPL_K1_PF_MB=0 UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras > gpurun_out/s2_b5_tail_nopf.json 2> gpurun_out/s2_b5_tail_nopf.err
for f in gpurun_out/s2_b5_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step']*1e3,1), {k:round(v*1e3,1) for k,v in d['roofline']['stages_ms'].items() if v is not None})
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
