"""A/B timing of K1 alone (the fused entry ubpl_warp_decode_k2, CUDA-graph replay, CUDA events): the current
library against an older build given as UBPL_OLD_LIB (round-1 ABI), and the current library under the
UBPL_K1_DBG masks (timing experiments: parts of the per-map work skipped, results void).
    python tools/k1_ab.py c2 c4"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import _lib, ops, synth  # noqa: E402

c_void_p, c_int, c_i64, c_float, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double
OLD_K2_SIG = [c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
              c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, c_int, c_float, c_float, c_int,
              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]


def timeit(fn, nbytes, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print("%-46s %7.1f us  %5.0f GB/s" % (label, us, nbytes / us / 1e3), flush=True)


def timeline(fn, stats):
    """UBPL_K1_DBG bit 16: the kernel's %globaltimer probe (warp_decode.cu tl_*), one eager launch."""
    big = 1 << 62
    stats.zero_()
    stats[8] = big; stats[10] = big; stats[12] = big
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    s = [int(x) for x in stats.tolist()]
    t0 = s[8]
    warps = max(s[2] and 1, 1)
    print("    timeline (us from the first CTA's start): last CTA start %.1f | first map landed %.1f .. %.1f | warps out of maps "
          "%.1f .. %.1f | last exit %.1f" % ((s[9] - t0) / 1e3, (s[10] - t0) / 1e3, (s[11] - t0) / 1e3, (s[12] - t0) / 1e3,
                                             (s[13] - t0) / 1e3, (s[14] - t0) / 1e3))
    live = max(s[19], 1)
    print("    share of the warps' live time: waiting for the staged copy %.3f, for a copy ticket %.3f, exhaustive decode %.3f, "
          "helping %.3f | maps %d, exhaustive %d, mean live time per warp %.1f us" %
          (s[15] / live, s[16] / live, s[17] / live, s[18] / live, s[2], s[0], live / 1e3 / 2072.0), flush=True)
    if len(s) > 32 + 2 * 16 * 148 and any(s[32:]):
        # UBPL_K1_DBG bit 32: per-warp records -> when did the warps / the CTAs run out of maps, and who were the last
        import numpy as np
        rec = np.array(s[32:32 + 2 * 16 * 148], dtype=np.int64).reshape(148, 16, 2)
        out = (rec[:, :, 0] - t0) / 1e3
        maps = rec[:, :, 1] & 0xffff
        slow = (rec[:, :, 1] >> 16) & 0xffff
        exh_us = (rec[:, :, 1] >> 32) / 1e3
        live = maps > 0
        o = out[live]
        print("    warps out of maps: percentiles 1/10/50/90/99/100 = %s us" % np.round(np.percentile(o, [1, 10, 50, 90, 99, 100]), 1).tolist())
        cta_last = np.where(live, out, 0).max(axis=1)
        cta_first = np.where(live, out, 1e9).min(axis=1)
        print("    CTAs: last warp out, percentiles 0/10/50/90/100 = %s us; first warp out = %s us" %
              (np.round(np.percentile(cta_last, [0, 10, 50, 90, 100]), 1).tolist(), np.round(np.percentile(cta_first, [0, 10, 50, 90, 100]), 1).tolist()))
        print("    maps per CTA min/median/max = %d/%d/%d; per warp min/median/max = %d/%d/%d" %
              (maps.sum(1).min(), np.median(maps.sum(1)), maps.sum(1).max(), maps[live].min(), np.median(maps[live]), maps[live].max()))
        late = np.argsort(-cta_last)[:8]
        for c in late:
            w = int(np.argmax(np.where(live[c], out[c], 0)))
            print("      late CTA %3d: last warp out %.1f us (warp %d: %d maps, %d exhaustive, %.1f us in them); CTA total %d maps, %d exhaustive, "
                  "%.1f us in them; first warp out %.1f" % (c, cta_last[c], w, maps[c, w], slow[c, w], exh_us[c, w], maps[c].sum(), slow[c].sum(),
                                                             exh_us[c].sum(), cta_first[c]))


def one(cfgname, c):
    d = synth.make_batch(B=c["B"], K=c["K"], J=c["J"], H=c["H"], W=c["W"], M=1, S=2, seed=1388, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [c["H"], c["W"]])
    nbytes = 4 * c["H"] * c["W"] * c["J"] * c["K"] * c["B"]
    t = d["teacher"][0]
    mode = 2 if c["select"] == "fixed" else 1
    old = os.environ.get("UBPL_OLD_LIB")
    if old and os.path.exists(old):
        L = ctypes.CDLL(old)
        L.ubpl_warp_decode_k2.argtypes = OLD_K2_SIG
        L.ubpl_warp_decode_k2_ws_bytes.restype = c_i64
        K, B, J, H, W = t.shape
        th = d["theta"].contiguous(); fl = d["flip"].to(torch.uint8).contiguous()
        o_idx = torch.empty(K, B, J, dtype=torch.int32, device="cuda"); o_max = torch.empty(K, B, J, device="cuda")
        o_xy = torch.empty(K, B, J, 2, device="cuda"); mean = torch.empty(B, J, 2, device="cuda")
        dist = torch.empty(B, J, dtype=torch.float64, device="cuda"); legal = torch.empty(B, J, dtype=torch.uint8, device="cuda")
        en = torch.empty(B, J, dtype=torch.uint8, device="cuda"); gate = torch.empty(B, J, device="cuda")
        wsb = int(L.ubpl_warp_decode_k2_ws_bytes(K, B, J))
        ws = torch.empty(wsb // 4, dtype=torch.int32, device="cuda")

        def run_old():
            rc = L.ubpl_warp_decode_k2(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2), K, B, J, H, W, th.data_ptr(),
                                       fl.data_ptr(), dec.data_ptr(), 0, o_idx.data_ptr(), o_max.data_ptr(), o_xy.data_ptr(),
                                       mode, 3.0, 256, 256, 4.0, 3.0, 2, mean.data_ptr(), dist.data_ptr(), legal.data_ptr(),
                                       en.data_ptr(), gate.data_ptr(), None, ws.data_ptr(), wsb, None,
                                       torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        timeit(run_old, nbytes, cfgname + " round-1 library (main + slow launch)")
    plan = None
    if os.environ.get("UBPL_AB_EMA", "0") != "0":           # the EMA of an HG2-sized model (8.4 M parameters) inside K1's launch
        ps = [torch.randn(n, device="cuda") for n in [1 << 20] * 8 + [40000]]
        es = [torch.randn_like(x) for x in ps]
        plan = ops.EmaPlan(ps, es)
    for cap in [int(x) for x in os.environ.get("UBPL_AB_CAPS", "8").split(",")]:
        os.environ["UBPL_K1_INFLIGHT"] = str(cap)
        for dbg in [int(x) for x in os.environ.get("UBPL_AB_MASKS", "0,4,1,3,11").split(",")]:
            os.environ["UBPL_K1_DBG"] = str(dbg)
            stats = torch.zeros(32 + 2 * 16 * 148 + 8, dtype=torch.int64, device="cuda") if (dbg & 16) else None
            run = lambda: ops.warp_decode_k2(t, d["theta"], d["flip"], dec, mode, S=2, distThrMax=3.0, stats=stats,  # noqa: E731
                                             ema=plan, alpha=0.99)
            timeit(run, nbytes, cfgname + " current, INFLIGHT=%d DBG=%d%s" % (cap, dbg, " + EMA in the tail" if plan else ""))
            if dbg & 16:
                timeline(run, stats)
    os.environ.pop("UBPL_K1_DBG", None)
    os.environ.pop("UBPL_K1_INFLIGHT", None)


for name in (sys.argv[1:] or ["c2"]):
    for Bov in [int(x) for x in os.environ.get("UBPL_AB_B", "0").split(",")]:      # UBPL_AB_B: batch sizes to sweep
        cfg = dict(bench.CONFIGS[name])
        if Bov:
            cfg["B"] = Bov
        one("%s@B%d" % (name, cfg["B"]), cfg)
