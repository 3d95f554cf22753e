#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest7.log
UBPL_AB_CAPS=0,6,8 UBPL_AB_MASKS=0,4 timeout 150 python tools/k1_ab.py c2 c4 > gpurun_out/r2_k1_ab3.log 2>&1; echo "ab rc=$?"; grep -v Warning gpurun_out/r2_k1_ab3.log
