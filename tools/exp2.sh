cd $GRAFT_REPO_ROOT
UBPL_AB_MASKS=48 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_perwarp_c2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s2_pytest1.log
timeout 600 python tools/k3_variants.py > gpurun_out/s2_k3_lean.log 2>&1
cat gpurun_out/s2_k3_lean.log | tail -12
cat gpurun_out/s2_k1_perwarp_c2.log | tail -20
