#!/bin/bash
# N GPUs: p2p_check + bench (c2 headline + c4 collective block); N=1 c4 first for the collective's denominator
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 300 python bench.py --config c4 --steps 100 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_bench18_c4_n1.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench18_c4_n1.json').read().strip().splitlines()[-1]);print('c4 n1',round(d['value']),round(d['ms_per_step']*1e3,1),{k:(round(v*1e3,1) if v else v) for k,v in d['roofline']['stages_ms'].items()})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | grep -v Warning | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench18_n$N.json 2> gpurun_out/r2_bench18_n$N.err; echo "bench n$N rc=$?"; tail -c 400 gpurun_out/r2_bench18_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench18_n$N.json').read().strip().splitlines()[-1])
print('n$N c2', round(d['value']), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']))
c=d.get('collective'); print('collective', round(c['value']), round(c['ms_per_step']*1e3,1), 'k1/k2/k3 us', round(c['k1_stage_us'],1), round(c['k2_stage_us'],1), round(c['k3_stage_us'],1))
PY
