"""Phase timing of the multi-rank (histogram-exchange) selector with its ranks emulated on one GPU:
python tools/select_phases_multi.py [R] [n]   (globaltimer stamps of rank 0)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import _lib, ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4352
J = 17
rng = np.random.default_rng(3)
for name, d in (("quarter-pixel grid (ties)", np.round(rng.gamma(2.0, 3.0, (R, n)) * 4) / 4), ("continuous", rng.gamma(2.0, 3.0, (R, n)))):
    dist = torch.as_tensor(d).cuda()
    legal = torch.as_tensor((rng.random((R, n)) < 0.9).astype(np.uint8)).cuda()
    stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
    _lib.call("ubpl_select_debug_stamps", stamps.data_ptr())
    for it in range(3):
        r = ops.select_quantile_emul(dist, legal, J, 0.0, 0.5, 1.0)
    torch.cuda.synchronize()
    s = stamps.cpu().tolist()
    c = s[63]
    print("%s: R = %d, n = %d: %d stamps, total %.1f us" % (name, R, n, c, (s[c - 1] - s[0]) / 1e3))
    print("  deltas (us):", [round((s[i + 1] - s[i]) / 1e3, 2) for i in range(c - 1)])
    _lib.call("ubpl_select_debug_stamps", None)
