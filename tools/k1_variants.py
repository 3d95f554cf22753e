"""Times K1 alone on a bench config (CUDA-graph replay, CUDA events): the plain entry (ubpl_warp_decode) and
the entry with the K2 epilogue (ubpl_warp_decode_k2, modes 1 and 2), for the staging knobs given in the
environment (UBPL_K1_EARLY / UBPL_K1_PF / UBPL_K1_WARPS).  UBPL_LIB=<other libubpl_b200.so> times the plain
entry of an older build.  `python tools/k1_variants.py [cfg]`."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import _lib, ops, synth  # noqa: E402

old = os.environ.get("UBPL_LIB")
if old:
    L = ctypes.CDLL(old)
    fn = L.ubpl_warp_decode
    fn.argtypes = _lib.SIGNATURES["ubpl_warp_decode"]
    fn.restype = ctypes.c_int
    _lib.lib().ubpl_warp_decode = fn

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
c = bench.CONFIGS[cfgname]
d = synth.make_batch(B=c["B"], K=c["K"], J=c["J"], H=c["H"], W=c["W"], M=1, S=1, seed=1388, device="cuda")
dec = ops.decode_coeffs(d["center"], d["scale"], [c["H"], c["W"]])
other = torch.empty(64 * 1024 * 1024, device="cuda")        # 256 MB: flush L2 between runs
nbytes = 4 * c["H"] * c["W"] * c["J"] * c["K"] * c["B"]
t = d["teacher"][0]


def run(mode):
    if mode == 0:
        return ops.warp_decode(t, d["theta"], d["flip"], dec)
    return ops.warp_decode_k2(t, d["theta"], d["flip"], dec, mode, S=2, distThrMax=3.0)


for mode in ((0,) if old else (0, 1, 2)):
    for it in range(3):
        r = run(mode)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r = run(mode)
    ts = []
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
    print("lib=%s mode=%d early=%s pf=%s: median %.1f us (%.0f GB/s), back-to-back %.1f us (%.0f GB/s)"
          % ("old" if old else "new", mode, os.environ.get("UBPL_K1_EARLY", "-"), os.environ.get("UBPL_K1_PF", "-"), ts[10] * 1e3,
             nbytes / ts[10] / 1e6, e0.elapsed_time(e1) / 50 * 1e3, nbytes / (e0.elapsed_time(e1) / 50) / 1e6), flush=True)
