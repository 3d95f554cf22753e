cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s2_pytest2.log
UBPL_AB_MASKS=0,48,4,11 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_v2_c2.log 2>&1
UBPL_K1_ENDGAME_PCT=0 UBPL_AB_MASKS=0,48 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_v2_c2_noeg.log 2>&1
UBPL_K1_ENDGAME_PCT=300 UBPL_AB_MASKS=0,48 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_v2_c2_eg300.log 2>&1
UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c4 c3 > gpurun_out/s2_k1_v2_c4.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/rw_micro tools/rw_micro.cu && /tmp/rw_micro > gpurun_out/s2_rw_micro.log 2>&1
timeout 600 python bench.py --no-extras > gpurun_out/s2_bench_c2_v2.json 2> gpurun_out/s2_bench_c2_v2.err
cat gpurun_out/s2_k1_v2_c2.log gpurun_out/s2_k1_v2_c2_noeg.log gpurun_out/s2_k1_v2_c2_eg300.log gpurun_out/s2_k1_v2_c4.log | grep -v "late CTA"
cat gpurun_out/s2_rw_micro.log
python -c "
import json; d=json.loads(open('gpurun_out/s2_bench_c2_v2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])"
