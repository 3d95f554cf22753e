"""ctypes binding of libubpl_b200.so (include/ubpl_b200.h).

The product path has NO fallback: if the shared library is missing, or an op is called with a
non-CUDA tensor, an exception is raised.  `build()` compiles the library in-tree with nvcc for
sm_100a (it cross-compiles without a GPU)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libubpl_b200.so")

c_void_p, c_int, c_i64, c_float, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> argtypes (restype is int unless noted); mirrors include/ubpl_b200.h one to one
SIGNATURES = {
    "ubpl_version": [],
    "ubpl_device_info": [c_void_p] * 4,
    "ubpl_warp_decode": [c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_warp_decode_k2": [c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                            c_void_p, c_void_p, c_void_p,
                            c_int, c_double, c_int, c_int, c_float, c_float, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p],
    "ubpl_warp_decode_k2_ema": [c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p,
                                c_int, c_double, c_int, c_int, c_float, c_float, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_i64, c_void_p, c_i64,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_float, c_void_p,
                                c_void_p],
    "ubpl_warp_decode_k2_ws_bytes": [c_int, c_int, c_int],
    "ubpl_warp_materialize": [c_void_p, c_i64, c_i64, c_void_p, c_i64, c_i64, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_view_dispersion": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "ubpl_mirror_w": [c_void_p, c_void_p, c_i64, c_int, c_void_p],
    "ubpl_coord_error": [c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p],
    "ubpl_pair_distance": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p],
    "ubpl_unc_normalize": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p],
    "ubpl_assess_dual": [c_void_p] * 5 + [c_int, c_int, c_int] + [c_void_p] * 9 + [c_void_p],
    "ubpl_dist_extrema": [c_void_p, c_i64, c_void_p, c_void_p],
    "ubpl_reliability": [c_void_p, c_void_p, c_i64, c_void_p, c_double, c_void_p, c_void_p, c_void_p],
    "ubpl_key_histogram": [c_void_p, c_i64, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "ubpl_select_descend": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "ubpl_select_apply": [c_void_p, c_i64, c_int, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_select_quantile_local": [c_void_p, c_void_p, c_i64, c_int, c_i64, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_select_quantile_dist": [c_void_p, c_void_p, c_i64, c_int, c_i64, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_mix_dists": [c_void_p, c_int, c_int, c_int, c_double] + [c_void_p] * 6 + [c_int, c_int, c_int] + [c_void_p] * 12 + [c_void_p],
    "ubpl_mix_unc": [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_double] + [c_void_p] * 10 + [c_void_p],
    "ubpl_select_quantile_fused": [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_double, c_double,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_float, c_float, c_int, c_float, c_void_p, c_void_p,
                                   c_int, c_void_p],
    "ubpl_select_quantile_emul": [c_void_p, c_void_p, c_int, c_void_p, c_i64, c_int, c_i64, c_double, c_double,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p],
    "ubpl_select_debug_stamps": [c_void_p],
    "ubpl_p2p_buffer_bytes": [c_int, c_i64],
    "ubpl_p2p_alloc": [c_int, c_i64, c_void_p],
    "ubpl_p2p_open": [c_void_p, c_int, c_int],
    "ubpl_p2p_close": [],
    "ubpl_p2p_ranks": [],
    "ubpl_p2p_status": [],
    "ubpl_nccl_unique_id": [c_void_p],
    "ubpl_nccl_init": [c_void_p, c_int, c_int],
    "ubpl_nccl_destroy": [],
    "ubpl_nccl_ranks": [],
    "ubpl_select_fixed": [c_void_p, c_void_p, c_i64, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_k2_view_fixed": [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_float, c_float, c_int,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_render_mse": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_i64, c_i64,
                        c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                        c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_render_mse_sum": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_i64, c_i64,
                            c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                            c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_render_targets": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p],
    "ubpl_dense_mse": [c_void_p, c_i64, c_i64, c_i64, c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_void_p, c_int, c_float,
                       c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_loss_finalize": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "ubpl_gate_prepare": [c_void_p, c_void_p, c_i64, c_int, c_int, c_float, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_scale": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p],
    "ubpl_view_kps": [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p, c_void_p],
    "ubpl_acc_pck": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_features_cov": [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ubpl_ema_multi_tensor": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_float, c_void_p, c_void_p],
    "ubpl_ema_flat": [c_void_p, c_void_p, c_i64, c_float, c_float, c_void_p],
}

RESTYPES = {"ubpl_warp_decode_k2_ws_bytes": c_i64, "ubpl_p2p_buffer_bytes": c_i64}      # everything else returns an int status code

_lib = None


class UbplError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu into csrc/libubpl_b200.so (nvcc, sm_100a, -lineinfo)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise UbplError("building libubpl_b200.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UbplError("libubpl_b200.so is missing at %s: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for the ubpl_b200 ops)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.ubpl_last_error.restype = ctypes.c_char_p
        L.ubpl_last_error.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the header and the library disagree
            fn.argtypes = args
            fn.restype = RESTYPES.get(name, c_int)
        _lib = L
    return _lib


# kernels launched per successful call (cudaMemsetAsync is not counted); bench.py's gpu_launches
LAUNCHES = {"ubpl_dist_extrema": 2, "ubpl_features_cov": 2, "ubpl_select_quantile_dist": 16, "ubpl_nccl_unique_id": 0,
            "ubpl_nccl_init": 0, "ubpl_nccl_destroy": 0, "ubpl_p2p_alloc": 0, "ubpl_p2p_open": 0, "ubpl_p2p_close": 0, "ubpl_select_debug_stamps": 0}
_launches = 0


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def call(name, *args):
    global _launches
    L = lib()
    rc = getattr(L, name)(*args)
    _launches += LAUNCHES.get(name, 1)
    if rc != 0:
        raise UbplError("%s failed (%d): %s" % (name, rc, L.ubpl_last_error().decode()))
