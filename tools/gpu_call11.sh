#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "multirank" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest11.log | cut -c1-300
timeout 200 python tools/k3_variants.py > gpurun_out/r2_k3_var.log 2>&1; echo "k3 rc=$?"; grep -v Warning gpurun_out/r2_k3_var.log
