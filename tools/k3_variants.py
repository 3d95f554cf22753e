"""Times ubpl_render_mse(_sum) alone on the c2 shapes (CUDA events, graph replay, L2 flushed) for kernel
experiments.  UBPL_LIB=<path to another libubpl_b200.so> times an older build through its plain entry."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import _lib, ops, synth  # noqa: E402

old = os.environ.get("UBPL_LIB")
if old:
    L = ctypes.CDLL(old)
    L.ubpl_last_error.restype = ctypes.c_char_p
    fn = L.ubpl_render_mse
    fn.argtypes = _lib.SIGNATURES["ubpl_render_mse"]
    fn.restype = ctypes.c_int
    _lib.lib().ubpl_render_mse = fn             # route the plain entry to the other build

d = synth.make_batch(B=256, K=1, J=14, M=1, S=2, device="cuda")
B, J = 256, 14
kps = (d["base_xy"] * 4 + 1).contiguous()
gate = (torch.rand(B, J, device="cuda") < 0.6).float()
w = torch.where(d["islabeled"], 0.0, 1.0).float()
other = torch.empty(64 * 1024 * 1024, device="cuda")        # 256 MB: flush L2 between runs
# (generic kernel's occupancy, item shared by a CTA, CTAs per SM override, lean kernel: 0 off / 4, 5, 6 = its occupancy)
VARIANTS = [(5, 1, 0, 0), (5, 1, 0, 5), (5, 1, 0, 4), (5, 1, 0, 6), (5, 1, 4, 5), (5, 1, 6, 5), (5, 1, 5, 4)]
for summ in ((False,) if old else (True,)):
    for occ, coop, ctas, fast in ([(6, 0, 0, 0)] if old else VARIANTS):
        os.environ["UBPL_K3_OCC"] = str(occ)
        os.environ["UBPL_K3_COOP"] = str(coop)
        os.environ["UBPL_K3_CTAS"] = str(ctas)
        os.environ["UBPL_K3_FAST"] = "1" if fast else "0"
        os.environ["UBPL_K3_FAST_OCC"] = str(fast or 5)
        for it in range(3):
            r = ops.render_mse(kps, gate, w, d["student"], 256, 256, want_summary=summ)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            r = ops.render_mse(kps, gate, w, d["student"], 256, 256, want_summary=summ)
        ts = []
        for it in range(20):
            other.fill_(1.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(50):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ts.sort()
        nbytes = 4 * 4096 * J * B * 5
        print("lib=%s summary=%s occ=%d coop=%d ctas/SM=%d lean=%d: median %.1f us (%.0f GB/s), back-to-back %.1f us (%.0f GB/s)"
              % ("old" if old else "new", summ, occ, coop, ctas, fast, ts[10] * 1e3, nbytes / ts[10] / 1e6, e0.elapsed_time(e1) / 50 * 1e3,
                 nbytes / (e0.elapsed_time(e1) / 50) / 1e6), flush=True)
