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


for cfgname in (sys.argv[1:] or ["c2"]):
    c = bench.CONFIGS[cfgname]
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
    for cap in [int(x) for x in os.environ.get("UBPL_AB_CAPS", "0").split(",")]:
        os.environ["UBPL_K1_INFLIGHT"] = str(cap)
        for dbg in [int(x) for x in os.environ.get("UBPL_AB_MASKS", "0,4,1,3,11").split(",")]:
            os.environ["UBPL_K1_DBG"] = str(dbg)
            timeit(lambda: ops.warp_decode_k2(t, d["theta"], d["flip"], dec, mode, S=2, distThrMax=3.0), nbytes,
                   cfgname + " current, INFLIGHT=%d DBG=%d" % (cap, dbg))
    os.environ.pop("UBPL_K1_DBG", None)
    os.environ.pop("UBPL_K1_INFLIGHT", None)
