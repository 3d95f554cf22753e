#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
UBPL_K1_INFLIGHT=4 timeout 120 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "k1_variants or exhaustive" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest(cap 4) rc=$?"; tail -3 gpurun_out/r2_pytest6.log
UBPL_AB_CAPS=0,3,4,5,6 UBPL_AB_MASKS=0,4 timeout 150 python tools/k1_ab.py c2 c4 c3 > gpurun_out/r2_k1_ab2.log 2>&1; echo "ab rc=$?"; grep -v Warning gpurun_out/r2_k1_ab2.log
