#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 tools/k1_micro 69632 > gpurun_out/r2_k1_micro_c4b.log 2>&1; echo "micro rc=$?"; grep -E "^(stage  14|spin|cap)" gpurun_out/r2_k1_micro_c4b.log
timeout 200 tools/k1_micro 28672 > gpurun_out/r2_k1_micro_c2b.log 2>&1; echo "micro rc=$?"; grep -E "^(stage  14|spin|cap)" gpurun_out/r2_k1_micro_c2b.log
UBPL_AB_CAPS=0,3,4,5,6,8 UBPL_AB_MASKS=0,4 timeout 600 python tools/k1_ab.py c2 c4 c3 > gpurun_out/r2_k1_ab2.log 2>&1; echo "ab rc=$?"; grep -v Warning gpurun_out/r2_k1_ab2.log
