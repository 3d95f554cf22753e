#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/k3_variants.py > gpurun_out/r2_k3_var.log 2>&1; echo "k3 rc=$?"; grep -v Warning gpurun_out/r2_k3_var.log
