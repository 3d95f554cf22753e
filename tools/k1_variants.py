"""Times K1 alone (ubpl_warp_decode_k2, mode 2) on a bench config for every staging variant:
UBPL_K1_EARLY x UBPL_K1_PF (x warps).  CUDA-graph replay, CUDA events, L2 flushed between runs.
`python tools/k1_variants.py [cfg]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import ops, synth  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
c = bench.CONFIGS[cfgname]
d = synth.make_batch(B=c["B"], K=c["K"], J=c["J"], H=c["H"], W=c["W"], M=1, S=1, seed=1388, device="cuda")
dec = ops.decode_coeffs(d["center"], d["scale"], [c["H"], c["W"]])
other = torch.empty(64 * 1024 * 1024, device="cuda")        # 256 MB: flush L2 between runs
nbytes = 4 * c["H"] * c["W"] * c["J"] * c["K"] * c["B"]
ref = None
for early in (0, 1):
    for pf in (0, 1):
        for warps in (0, 12):
            os.environ["UBPL_K1_EARLY"], os.environ["UBPL_K1_PF"], os.environ["UBPL_K1_WARPS"] = str(early), str(pf), str(warps)
            stats = torch.zeros(4, dtype=torch.int64, device="cuda")
            for it in range(3):
                r = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 2, S=2, distThrMax=3.0, stats=stats)
            torch.cuda.synchronize()
            st = stats.tolist()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                r = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 2, S=2, distThrMax=3.0)
            ts, tb = [], []
            for it in range(20):
                other.fill_(1.0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(50):                                   # back to back (what the step sees)
                g.replay()
            e1.record(); torch.cuda.synchronize()
            ts.sort()
            sig = (r["idx"].clone(), r["enable"].clone())
            same = True if ref is None else (torch.equal(sig[0], ref[0]) and torch.equal(sig[1], ref[1]))
            if ref is None:
                ref = sig
            print("early=%d pf=%d warps=%2d: median %.1f us (%.0f GB/s), back-to-back %.1f us (%.0f GB/s); slow %d miss %d of %d maps; same=%s"
                  % (early, pf, warps, ts[10] * 1e3, nbytes / ts[10] / 1e6, e0.elapsed_time(e1) / 50 * 1e3,
                     nbytes / (e0.elapsed_time(e1) / 50) / 1e6, st[0] // 3, st[3] // 3, st[2] // 3, same), flush=True)
