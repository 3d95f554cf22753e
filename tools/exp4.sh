cd $GRAFT_REPO_ROOT
UBPL_AB_MASKS=0,16,4 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 c4 c3 > gpurun_out/s2_k1_v3.log 2>&1
grep -v "late CTA" gpurun_out/s2_k1_v3.log
for ov in k1 tail; do
  UBPL_BENCH_OVERLAP_EMA=$ov timeout 600 python bench.py --no-extras > gpurun_out/s2_b4_$ov.json 2> gpurun_out/s2_b4_$ov.err
done
for tp in 100 400 800; do
  UBPL_K1_TRIGGER_PCT=$tp UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras > gpurun_out/s2_b4_tail_t$tp.json 2> gpurun_out/s2_b4_tail_t$tp.err
done
UBPL_EMA_CTAS=4 UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras > gpurun_out/s2_b4_tail_c4.json 2> gpurun_out/s2_b4_tail_c4.err
UBPL_BENCH_OVERLAP_EMA=tail timeout 600 python bench.py --no-extras --config c4 > gpurun_out/s2_b4_c4_tail.json 2> gpurun_out/s2_b4_c4_tail.err
UBPL_BENCH_OVERLAP_EMA=k1 timeout 600 python bench.py --no-extras --config c4 > gpurun_out/s2_b4_c4_k1.json 2> gpurun_out/s2_b4_c4_k1.err
for f in gpurun_out/s2_b4_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step']*1e3,1), {k:round(v*1e3,1) for k,v in d['roofline']['stages_ms'].items() if v is not None})
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
