"""Multi-GPU check of the peer-memory selector (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/p2p_check.py
Every rank owns a shard; the fused P2P selection must equal (bit for bit) the single-GPU selection over the
concatenated shards, and the NCCL histogram selector.  Also times both selectors (CUDA graph replay)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import dist as ud, ops  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
group = td.group.WORLD
ok = ud.init_p2p(group, max_items=8192)
if rank == 0:
    print("init_p2p:", ok, flush=True)
if not ok:
    td.destroy_process_group()
    sys.exit(3)
ud.init_nccl(group)
fails = 0
for trial, (n, J) in enumerate(((7, 7), (4352, 17), (4352, 17), (8192, 16), (1000, 10))):
    rng = np.random.default_rng(100 + trial)
    dist_all = np.round(rng.gamma(2.0, 3.0, n * world) * 4) / 4
    dist_all[rng.random(n * world) < 0.2] = 999.0
    legal_all = (rng.random(n * world) < 0.9).astype(np.uint8)
    kps_all = rng.uniform(-5, 261, (n * world, 2)).astype(np.float32)
    sl = slice(rank * n, (rank + 1) * n)
    dl, ll, kl = (torch.as_tensor(x[sl]).to(dev) for x in (dist_all, legal_all, kps_all))
    for pct in (0.5, 0.1, 0.99):
        k = int((n * world - 1) * pct)
        gate = (kl, 2, 256, 256, 4.0, 3.0, 0.7)
        a = ops.select_quantile_fused(dl, ll, J, k, 0.0, 1.0, gate=gate, p2p=True)
        full = ops.select_quantile_fused(torch.as_tensor(dist_all).to(dev), torch.as_tensor(legal_all).to(dev), J, k, 0.0, 1.0,
                                         gate=(torch.as_tensor(kps_all).to(dev), 2, 256, 256, 4.0, 3.0, 0.7))
        b = ops.select_quantile_nccl(dl, ll.double(), J, k, 0.0, 1.0)
        torch.cuda.synchronize()
        good = (float(a["thr"]) == float(full["thr"]) == float(b["thr"]) and torch.equal(a["enable"], full["enable"][sl])
                and torch.equal(a["enable"], b["enable"]) and torch.equal(a["reliability"], full["reliability"][sl])
                and torch.equal(a["gate"], full["gate"][sl]) and torch.equal(a["ext"], full["ext"]))
        if not good:
            fails += 1
            print("rank %d MISMATCH n=%d pct=%.2f thr %r %r %r" % (rank, n, pct, float(a["thr"]), float(full["thr"]), float(b["thr"])), flush=True)
assert ud.p2p_status() == 0, "a peer timed out"
t = torch.tensor([fails], device=dev)
td.all_reduce(t)
if rank == 0:
    print("p2p selector mismatches over all ranks:", int(t), flush=True)

# timing: graph replay of the P2P selector vs the eager NCCL selector (n = 4352 per rank, the c4 shard)
n, J = 4352, 17
rng = np.random.default_rng(7 + rank)
dl = torch.as_tensor(np.round(rng.gamma(2.0, 3.0, n) * 4) / 4).to(dev)
ll = torch.as_tensor((rng.random(n) < 0.9).astype(np.uint8)).to(dev)
kl = torch.as_tensor(rng.uniform(-5, 261, (n, 2)).astype(np.float32)).to(dev)
k = int((n * world - 1) * 0.5)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        ops.select_quantile_fused(dl, ll, J, k, 0.0, 1.0, gate=(kl, 2, 256, 256, 4.0, 3.0, 1.0), p2p=True)
torch.cuda.synchronize()
td.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    r = ops.select_quantile_fused(dl, ll, J, k, 0.0, 1.0, gate=(kl, 2, 256, 256, 4.0, 3.0, 1.0), p2p=True)
torch.cuda.synchronize()
td.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    g.replay()
e1.record()
torch.cuda.synchronize()
p2p_us = e0.elapsed_time(e1) / 200 * 1e3
lld = ll.double()
for _ in range(5):
    ops.select_quantile_nccl(dl, lld, J, k, 0.0, 1.0)
torch.cuda.synchronize()
td.barrier()
e0.record()
for _ in range(100):
    ops.select_quantile_nccl(dl, lld, J, k, 0.0, 1.0)
e1.record()
torch.cuda.synchronize()
nccl_us = e0.elapsed_time(e1) / 100 * 1e3
assert ud.p2p_status() == 0
if rank == 0:
    print("selector, %d ranks x %d items: peer-memory kernel (graph) %.1f us, NCCL histogram path (eager) %.1f us" % (world, n, p2p_us, nccl_us), flush=True)
ud.destroy_p2p()
ud.destroy_nccl()
td.destroy_process_group()
sys.exit(1 if int(t) else 0)
