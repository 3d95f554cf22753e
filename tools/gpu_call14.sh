#!/bin/bash
# 2 GPUs: the real peer-memory selector + the bench's collective block
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 400 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_parity.py -m gpu -x -q -k "p2p or dense or pseudo3 or dist_loss or mt2" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest14.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | grep -v Warning | tail -5
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench14_n2.json 2> gpurun_out/r2_bench14_n2.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r2_bench14_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench14_n2.json').read().strip().splitlines()[-1])
print('n2 c2', round(d['value']), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']))
print('collective', json.dumps(d.get('collective'), indent=0))
PY
timeout 300 python bench.py --config c4 --steps 100 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_bench14_c4_n1.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench14_c4_n1.json').read().strip().splitlines()[-1]);print('c4 n1',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
