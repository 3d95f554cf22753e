#!/bin/bash
# The GPU-side measurement recipe of a round (run through gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02'
# parity tests, the bench lines of every single-GPU config, the launch list of the default bench and ONE
# `ncu --set full` capture of the step's kernels (each only after the plain command has exited 0).
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/<round>/.
tag=${1:-round}
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; echo "bench c2 rc=$?"
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
for f in gpurun_out/${tag}_bench_c?.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1], round(d["value"]), "samples/s", round(d["ms_per_step"] * 1e3, 1), "us/step",
          {k: round(v * 1e3, 1) for k, v in r["stages_ms"].items() if v is not None}, "frac", round(r["frac"], 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
# launch list (shares of the step; serialised and cold under ncu) and one full capture of the step's kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ubpl|warp_decode|render_mse|ema_|dense_mse|select_|k2_" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches_ncu.log 2>&1; echo "launch list rc=$?"
python tools/prof_step.py c2 4 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_decode_kernel|render_mse_kernel" -s 4 -c 2 \
    -o gpurun_out/${tag}_k1_k3_full python tools/prof_step.py c2 4 > gpurun_out/${tag}_prof_ncu.log 2>&1; echo "ncu rc=$?"
