#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 tools/k1_micro 28672 > gpurun_out/r2_k1_micro_c2.log 2>&1; echo "micro rc=$?"; cat gpurun_out/r2_k1_micro_c2.log
timeout 200 tools/k1_micro 69632 > gpurun_out/r2_k1_micro_c4.log 2>&1; echo "micro rc=$?"; cat gpurun_out/r2_k1_micro_c4.log
