"""Phase timing of the fused quantile selector (globaltimer stamps written by thread 0): python tools/select_phases.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4352
J = 17
rng = np.random.default_rng(3)
dist = torch.as_tensor(np.round(rng.gamma(2.0, 3.0, n) * 4) / 4).cuda()
legal = torch.as_tensor((rng.random(n) < 0.9).astype(np.uint8)).cuda()
kps = torch.as_tensor(rng.uniform(-5, 261, (n, 2)).astype(np.float32)).cuda()
stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.call("ubpl_select_debug_stamps", stamps.data_ptr())
for it in range(3):
    ops.select_quantile_fused(dist, legal, J, (n - 1) // 2, 0.0, 1.0, gate=(kps, 2, 256, 256, 4.0, 3.0, 1.0))
torch.cuda.synchronize()
s = stamps.cpu().tolist()
c = s[63]
print("n =", n, "stamps", c, "total %.1f us" % ((s[c - 1] - s[0]) / 1e3))
print("deltas (us):", [round((s[i + 1] - s[i]) / 1e3, 2) for i in range(c - 1)])
_lib.call("ubpl_select_debug_stamps", None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.select_quantile_fused(dist, legal, J, (n - 1) // 2, 0.0, 1.0, gate=(kps, 2, 256, 256, 4.0, 3.0, 1.0))
e0.record()
for it in range(100):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graph replay: %.1f us per launch" % (e0.elapsed_time(e1) / 100 * 1e3))
