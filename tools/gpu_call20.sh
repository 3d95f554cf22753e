#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest20.log | cut -c1-300
bash tools/sanitize.sh memcheck
