import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ubpl_b200
from ubpl_b200 import ops, _lib
for (R, n) in [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(3, 20000)]:
    rng = np.random.default_rng(R * 1000 + n)
    d = np.round(rng.gamma(2.0, 2.0, (R, n)) * 4) / 4
    legal = rng.random((R, n)) > 0.1
    d[~legal] = 999.0
    dist = torch.from_numpy(d).cuda(); leg = torch.from_numpy(legal.astype(np.uint8)).cuda()
    stamps = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
    _lib.call("ubpl_select_debug_stamps", stamps.data_ptr())
    r = ops.select_quantile_emul(dist, leg, 14, 0.0, 0.5, 1.0)
    torch.cuda.synchronize()
    s = stamps.cpu().tolist(); c = s[63]
    one = ops.select_quantile_fused(dist.reshape(-1), leg.reshape(-1), 14, int((R * n - 1) * 0.5), 0.0, 1.0)
    print(R, n, "status", r["status"].tolist(), "thr", r["thr"].tolist(), "want", float(one["thr"]), "stamps", c,
          [round((s[i + 1] - s[i]) / 1e3, 1) for i in range(max(c - 1, 0))], flush=True)
    _lib.call("ubpl_select_debug_stamps", None)
