#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "multirank" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest17.log | cut -c1-400
for i in 1 2 3; do timeout 60 python tools/sel_debug.py 3,16384 3,20000 8,4352 16,3000 2>&1 | grep -v Warning | cut -c1-100; done
for R in 2 8; do timeout 60 python tools/select_phases_multi.py $R 4352 2>&1 | grep -v Warning; done
