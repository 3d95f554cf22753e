"""Times ubpl_render_mse alone on the c2 shapes (CUDA events, graph replay) for kernel experiments."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ubpl_b200
from ubpl_b200 import ops, synth

d = synth.make_batch(B=256, K=1, J=14, M=1, S=2, device="cuda")
B, J = 256, 14
kps = (d["base_xy"] * 4 + 1).contiguous()
gate = (torch.rand(B, J, device="cuda") < 0.6).float()
w = torch.where(d["islabeled"], 0.0, 1.0).float()
other = torch.empty(64 * 1024 * 1024, device="cuda")        # 256 MB: flush L2 between runs
for tgt in (True, False):
    for it in range(3):
        r = ops.render_mse(kps, gate, w, d["student"], 256, 256, want_target=tgt)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r = ops.render_mse(kps, gate, w, d["student"], 256, 256, want_target=tgt)
    ts = []
    for it in range(20):
        other.fill_(1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    nbytes = 4 * 4096 * J * B * (2 * 2 + (1 if tgt else 0))
    print("store=%s target=%s median %.1f us  %.0f GB/s" % (os.environ.get("UBPL_K3_STORE", "0"), tgt, ts[10] * 1e3, nbytes / ts[10] / 1e6))
