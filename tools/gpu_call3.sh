#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
UBPL_OLD_LIB=tools/libubpl_b200_r1.so timeout 600 python tools/k1_ab.py c2 c4 > gpurun_out/r2_k1_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_k1_ab.log | grep -v Warning
