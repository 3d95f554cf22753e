#!/bin/bash
# compute-sanitizer over the hot path, ONE tool per invocation (B200_PROFILING.md: one sanitizer tool per gpurun call):
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'      (memcheck | racecheck | synccheck | initcheck)
# The log goes to gpurun_out/sanitize_<tool>.log; copies of the runs of record live under profiles/r02/.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
TOOL=${1:-memcheck}
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 800 compute-sanitizer --tool "$TOOL" --print-limit 20 python tools/sanitize_target.py > "gpurun_out/sanitize_$TOOL.log" 2>&1
echo "sanitizer rc=$?"
grep -v Warning "gpurun_out/sanitize_$TOOL.log" | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|step M|emulated|done|Error|error" | head -40
