cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest5.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/s2_pytest5.log
UBPL_AB_EMA=0 UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 c4 c3 > gpurun_out/s2_k1_v6.log 2>&1
UBPL_AB_EMA=1 UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 c4 c3 >> gpurun_out/s2_k1_v6.log 2>&1
grep -v "share of" gpurun_out/s2_k1_v6.log
for ov in k1 tail; do
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras > gpurun_out/s2_b8_$ov.json 2> gpurun_out/s2_b8_$ov.err
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras --config c3 > gpurun_out/s2_b8_c3_$ov.json 2> gpurun_out/s2_b8_c3_$ov.err
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras --config c4 > gpurun_out/s2_b8_c4_$ov.json 2> gpurun_out/s2_b8_c4_$ov.err
done
for f in gpurun_out/s2_b8_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step']*1e3,1), {k:round(v*1e3,1) for k,v in d['roofline']['stages_ms'].items() if v is not None})
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
