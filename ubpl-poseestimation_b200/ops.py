"""Tensor-level wrappers over the C ABI (include/ubpl_b200.h).  torch is used for device memory,
streams and (in dist.py) the NCCL plumbing only; every computation is a kernel of libubpl_b200.so.
All inputs must be CUDA tensors: there is no CPU path."""
import os

import torch

from . import _lib

_f32, _f64 = torch.float32, torch.float64
PF_CAP_MB = 64          # cap of K1's tail prefetch of the next kernel's input (MB)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.UbplError("ubpl_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                                 % t.device.type)


def _p(t):
    return None if t is None else t.data_ptr()


def _inner_contig(t):
    """Heat-map planes must be contiguous in (H, W); outer dims may have any stride."""
    W = t.shape[-1]
    if t.stride(-1) == 1 and t.stride(-2) == W:
        return t
    return t.contiguous()


# -------------------------------------------------------------------------------------------------
# K1
# -------------------------------------------------------------------------------------------------
def decode_coeffs(center, scale, res):
    """[B,4] float64 (a00, a02, a11, a12): the inverse of utils/udaap/transforms.py:119-130
    `get_transform(center, scale, res)` exactly as `transform(..., invert=1)` (:151-158) obtains
    it -- the arithmetic dtype follows the dtype of the `scale` tensor (float64 tensor -> float64,
    anything else -> float32), then a00 = 1/t00, a02 = -(t02 * a00) in float64 (np.linalg.inv)."""
    scale = torch.as_tensor(scale)
    center = torch.as_tensor(center)
    ft = _f64 if scale.dtype == _f64 else _f32
    sc = scale.reshape(-1).to(ft)
    c = center.reshape(-1, 2).to(_f64).to(ft)
    h = sc * 200
    # `python_float / tensor` is reciprocal(tensor) * python_float in torch (Tensor.__rtruediv__), which the
    # reference's get_transform goes through for every term: one ulp from the true quotient for ~24 % of the
    # (centre, scale) pairs when the numerator is not a power of two
    rh = torch.reciprocal(h)
    t00 = rh * float(res[1])
    t11 = rh * float(res[0])
    t02 = res[1] * (rh * (-c[:, 0]) + 0.5)
    t12 = res[0] * (rh * (-c[:, 1]) + 0.5)
    t00, t11, t02, t12 = t00.to(_f64), t11.to(_f64), t02.to(_f64), t12.to(_f64)
    a00 = 1.0 / t00
    a11 = 1.0 / t11
    return torch.stack([a00, -(t02 * a00), a11, -(t12 * a11)], -1).contiguous()


def _swap_perm(swap_perm, J, dev):
    """int32[J] device table of flip_back's channel exchange (utils/udaap/transforms.py:20-57), or None."""
    if swap_perm is None:
        return None
    t = torch.as_tensor(swap_perm).reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    if t.numel() != J:
        raise _lib.UbplError("swap_perm must have one entry per joint (%d), got %d" % (J, t.numel()))
    return t


def swap_perm_from_pairs(pairs, J):
    """The source-channel table equivalent to flip_back's SEQUENTIAL pair exchanges
    (utils/udaap/transforms.py:51-54): after the loop, out[:, j] = mirrored[:, perm[j]].  The '_300w' table of
    the reference names channel 26 twice, so the result is a general permutation, not an involution."""
    perm = list(range(J))
    for a, b in pairs:
        perm[a], perm[b] = perm[b], perm[a]
    return torch.tensor(perm, dtype=torch.int32)


def warp_decode(maps, theta=None, flip=None, dec=None, refine=0, stats=None, want_idx=True, want_hm=False,
                swap_perm=None):
    """Fused back-warp + flip + arg-max decode (K1).  maps [V,B,J,H,W] or [B,J,H,W] (V=1);
    theta [V,B,2,3] (None = plain decode of the raw maps), flip [V,B] bool/uint8, dec [B,4] float64
    from decode_coeffs (None = heat-map coordinates), swap_perm [J] (None = no left/right exchange, the
    reference's live path).  Returns dict(idx, max, xy[, hm_xy]) shaped like the leading dims of `maps`."""
    _need_cuda(maps, theta, flip, dec, stats)
    if maps.dtype != _f32:
        raise _lib.UbplError("heat-maps must be float32")
    squeeze = maps.dim() == 4
    if squeeze:
        maps = maps.unsqueeze(0)
    maps = _inner_contig(maps)
    V, B, J, H, W = maps.shape
    dev = maps.device
    if theta is not None:
        theta = theta.reshape(V, B, 2, 3).to(_f32).contiguous()
    if flip is not None:
        flip = flip.reshape(V, B).to(torch.uint8).contiguous()
    if dec is not None:
        dec = dec.reshape(B, 4).to(_f64).contiguous()
    perm = _swap_perm(swap_perm, J, dev)
    out_idx = torch.empty(V, B, J, dtype=torch.int32, device=dev) if want_idx else None
    out_max = torch.empty(V, B, J, dtype=_f32, device=dev)
    out_xy = torch.empty(V, B, J, 2, dtype=_f32, device=dev)
    out_hm = torch.empty(V, B, J, 2, dtype=_f32, device=dev) if want_hm else None
    ws = torch.empty(4, dtype=torch.int32, device=dev)          # the launch's private work-claim counter
    _lib.call("ubpl_warp_decode", maps.data_ptr(), maps.stride(0), maps.stride(1), maps.stride(2), V, B, J, H, W,
              _p(theta), _p(flip), _p(perm), _p(dec), 1 if theta is not None else 0, int(refine),
              _p(out_idx), _p(out_max), _p(out_xy), _p(out_hm), _p(stats), ws.data_ptr(), _stream())
    res = dict(idx=out_idx, max=out_max, xy=out_xy, hm_xy=out_hm)
    if squeeze:
        res = {k: (v[0] if v is not None else None) for k, v in res.items()}
    return res


def warp_decode_k2(maps, theta, flip, dec, mode, S=1, img_h=256, img_w=256, stride=4.0, sigma=3.0, distThrMax=1.0,
                   refine=0, stats=None, want_idx=True, swap_perm=None, prefetch=None, ema=None, alpha=None,
                   alpha_from_device=False):
    """K1 with the per-joint part of K2 fused into its epilogue (maps [V,B,J,H,W], V <= 32).
    One teacher (V = K views):
      mode 1: + mean [B,J,2], dist [B,J] f64 (999 = illegal), legal [B,J]   (utils/evaluation.py:44-54)
      mode 2: + enable, gate = enable * visibility, counts [J+1], count = S * #(gate > 0)
              (utils/business.py:237-261,375-376, utils/process.py:262-268, utils/losses.py:29)
    Two teachers (V = 2K maps, teacher-major; theta/flip given per map):
      mode 3: mean = float32 ensemble coordinate, dist = extDist, legal, zero_div (utils/business.py:109-161)
      mode 4: + the fixed rule on extDist, gate and counts as in mode 2.
    prefetch: a contiguous tensor the next kernel reads (the student maps of K3): warps that run out of maps pull it
    into L2 while the last maps finish.
    ema (an EmaPlan) + alpha: K4 inside the same launch (ubpl_warp_decode_k2_ema) -- the warps that have run out of maps
    do the mean-teacher EMA while the last maps are decoded; alpha_from_device=True reads {alpha, 1-alpha} from the
    plan's device buffer (EmaPlan.set_alpha), which is what a captured CUDA graph needs.
    The returned `status` (int32[1], device) is non-zero when a hand-off word of
    the epilogue never arrived -- check it with check_status() at a point where a sync is acceptable."""
    _need_cuda(maps, theta, flip, dec, stats)
    if maps.dtype != _f32:
        raise _lib.UbplError("heat-maps must be float32")
    maps = _inner_contig(maps)
    K, B, J, H, W = maps.shape
    if not 1 <= K <= 32:
        raise _lib.UbplError("warp_decode_k2 supports 1..32 views")
    dev = maps.device
    theta = theta.reshape(K, B, 2, 3).to(_f32).contiguous()
    flip = None if flip is None else flip.reshape(K, B).to(torch.uint8).contiguous()
    dec = None if dec is None else dec.reshape(B, 4).to(_f64).contiguous()
    perm = _swap_perm(swap_perm, J, dev)
    out_idx = torch.empty(K, B, J, dtype=torch.int32, device=dev) if want_idx else None
    out_max = torch.empty(K, B, J, dtype=_f32, device=dev)
    out_xy = torch.empty(K, B, J, 2, dtype=_f32, device=dev)
    mean = torch.empty(B, J, 2, dtype=_f32, device=dev)
    dist = torch.empty(B, J, dtype=_f64, device=dev)
    legal = torch.empty(B, J, dtype=torch.uint8, device=dev)
    enable = torch.empty(B, J, dtype=torch.uint8, device=dev) if mode in (2, 4) else None
    gate = torch.empty(B, J, dtype=_f32, device=dev) if mode in (2, 4) else None
    ws_bytes = int(_lib.lib().ubpl_warp_decode_k2_ws_bytes(K, B, J))
    ws = torch.empty(ws_bytes // 4, dtype=torch.int32, device=dev)
    pf_ptr, pf_bytes = None, 0
    if prefetch is not None and prefetch.is_contiguous() and prefetch.data_ptr() % 16 == 0:
        # at most PF_CAP_MB: what is pulled in must still be in the 126 MB L2 when the next kernel reads it
        cap = int(os.environ.get("UBPL_K1_PF_MB", PF_CAP_MB)) << 20
        pf_ptr, pf_bytes = prefetch.data_ptr(), min(prefetch.numel() * prefetch.element_size(), cap)
    args = (maps.data_ptr(), maps.stride(0), maps.stride(1), maps.stride(2), K, B, J, H, W,
            theta.data_ptr(), _p(flip), _p(perm), _p(dec), int(refine), _p(out_idx), out_max.data_ptr(),
            out_xy.data_ptr(), int(mode), float(distThrMax), int(img_h), int(img_w), float(stride), float(sigma),
            int(S), mean.data_ptr(), dist.data_ptr(), legal.data_ptr(), _p(enable), _p(gate), _p(stats),
            ws.data_ptr(), ws_bytes, pf_ptr, int(pf_bytes))
    if ema is not None:
        _lib.call("ubpl_warp_decode_k2_ema", *args, *ema.launch_args(alpha, alpha_from_device, pieces=True), _stream())
    else:
        _lib.call("ubpl_warp_decode_k2", *args, _stream())
    return dict(idx=out_idx, max=out_max, xy=out_xy, mean=mean, dist=dist, legal=legal, enable=enable, gate=gate,
                counts=ws[128:128 + J + 1], count=ws[128 + J + 1:128 + J + 2], zero_div=ws[34:35], status=ws[35:36],
                ws=ws)


def check_status(status, what="ubpl_warp_decode_k2"):
    """Raises when a device status word is set (one D2H sync): the K2 epilogue's hand-off timed out, the outputs of
    that launch are void."""
    if status is not None and int(status.reshape(-1)[0].item()) != 0:
        raise _lib.UbplError("%s: device status %d -- a hand-off word never arrived; the results of that launch are "
                             "void" % (what, int(status.reshape(-1)[0].item())))


def warp_materialize(heatmap, warpmat, isflip, swap_perm=None):
    """AugmentUtils.affine_back2 (utils/augment.py:37-47) as one kernel; returns a new tensor.  swap_perm [C]
    adds flip_back's left/right channel exchange for the flipped samples (utils/udaap/transforms.py:20-57)."""
    _need_cuda(heatmap, warpmat, isflip)
    x = _inner_contig(heatmap.to(_f32))
    N, C, H, W = x.shape
    out = torch.empty(N, C, H, W, dtype=_f32, device=x.device)
    theta = warpmat.reshape(N, 2, 3).to(_f32).contiguous()
    flip = None if isflip is None else torch.as_tensor(isflip, device=x.device).reshape(N).to(torch.uint8).contiguous()
    perm = _swap_perm(swap_perm, C, x.device)
    _lib.call("ubpl_warp_materialize", x.data_ptr(), x.stride(0), x.stride(1), out.data_ptr(), out.stride(0),
              out.stride(1), N, C, H, W, theta.data_ptr(), _p(flip), _p(perm), _stream())
    return out


def mirror_w(x):
    """Exact mirror of the last axis (utils/augment.py:247-252)."""
    _need_cuda(x)
    x = x.to(_f32).contiguous()
    out = torch.empty_like(x)
    W = x.shape[-1]
    _lib.call("ubpl_mirror_w", x.data_ptr(), out.data_ptr(), x.numel() // W, W, _stream())
    return out


# -------------------------------------------------------------------------------------------------
# K2
# -------------------------------------------------------------------------------------------------
def view_dispersion(preds, sentinel_illegal=False, mean_in=None):
    """preds [K,B,J,2] float32 -> dict(mean [B,J,2] f32, dist [B,J] f64, unc32 [B,J] f32,
    legal [B,J] uint8, max_bits uint32 scalar)   (utils/evaluation.py:40-55)."""
    _need_cuda(preds)
    preds = preds.to(_f32).contiguous()
    K, B, J, _ = preds.shape
    dev = preds.device
    mean = torch.empty(B, J, 2, dtype=_f32, device=dev)
    dist = torch.empty(B, J, dtype=_f64, device=dev)
    unc32 = torch.empty(B, J, dtype=_f32, device=dev)
    legal = torch.empty(B, J, dtype=torch.uint8, device=dev)
    max_bits = torch.zeros(1, dtype=torch.int32, device=dev)
    if mean_in is not None:
        _need_cuda(mean_in)
        mean_in = mean_in.reshape(B, J, 2).to(_f32).contiguous()
    _lib.call("ubpl_view_dispersion", preds.data_ptr(), _p(mean_in), K, B, J, mean.data_ptr(), dist.data_ptr(), unc32.data_ptr(),
              legal.data_ptr(), max_bits.data_ptr(), 1 if sentinel_illegal else 0, _stream())
    return dict(mean=mean, dist=dist, unc32=unc32, legal=legal, max_bits=max_bits)


def unc_normalize(unc32, max_bits):
    """unc/unc.max() and exp(-unc)  (utils/evaluation.py:56-57)."""
    _need_cuda(unc32, max_bits)
    unc = torch.empty_like(unc32)
    uncW = torch.empty_like(unc32)
    _lib.call("ubpl_unc_normalize", unc32.data_ptr(), max_bits.data_ptr(), unc32.numel(), unc.data_ptr(),
              uncW.data_ptr(), _stream())
    return unc, uncW


def assess_dual(p1, p2, pmean, aug1, aug2):  # pmean None = bus.preds_mean(p1, p2)
    """utils/business.py:109-161 in array form; all outputs float64 [B,J] (coord [B,J,2])."""
    _need_cuda(p1, p2, pmean, aug1, aug2)
    aug1 = aug1.to(_f32).contiguous()
    aug2 = aug2.to(_f32).contiguous()
    K, B, J, _ = aug1.shape
    dev = aug1.device
    p1, p2 = (t.to(_f32).contiguous() for t in (p1, p2))
    pmean = None if pmean is None else pmean.to(_f32).contiguous()
    o = {k: torch.empty(B, J, dtype=_f64, device=dev) for k in ("legal", "intDist1", "intDist2", "extDist", "w1", "w2")}
    o["coord"] = torch.empty(B, J, 2, dtype=_f64, device=dev)
    o["coord32"] = torch.empty(B, J, 2, dtype=_f32, device=dev)
    o["zero_div"] = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("ubpl_assess_dual", p1.data_ptr(), p2.data_ptr(), _p(pmean), aug1.data_ptr(), aug2.data_ptr(),
              K, B, J, o["legal"].data_ptr(), o["intDist1"].data_ptr(), o["intDist2"].data_ptr(),
              o["extDist"].data_ptr(), o["w1"].data_ptr(), o["w2"].data_ptr(), o["coord"].data_ptr(),
              o["coord32"].data_ptr(), o["zero_div"].data_ptr(), _stream())
    return o


def coord_error(pred, gt, pck_ref, pck_thr):
    """pred [n_sets,B,J,2] (any float dtype) vs gt [B,J,>=2] float32 -> (err float64, acc int32), both
    [n_sets,B,J]  (utils/business.py:37-40, utils/evaluation.py:78-89)."""
    _need_cuda(pred, gt)
    pred = pred.to(_f64).contiguous()
    gt = gt.to(_f32).contiguous()
    n_sets, B, J, _ = pred.shape
    err = torch.empty(n_sets, B, J, dtype=_f64, device=pred.device)
    acc = torch.empty(n_sets, B, J, dtype=torch.int32, device=pred.device)
    _lib.call("ubpl_coord_error", pred.data_ptr(), gt.data_ptr(), gt.shape[-1], n_sets, B, J, int(pck_ref[0]),
              int(pck_ref[1]), float(pck_thr), err.data_ptr(), acc.data_ptr(), _stream())
    return err, acc


class _CudaSelectBackend:
    """The five device primitives of the quantile selector (K2 kernels of libubpl_b200.so)."""

    def prepare(self, dist, legal):
        _need_cuda(dist, legal)
        return dist.reshape(-1).to(_f64).contiguous(), legal.reshape(-1).to(_f64).contiguous()

    def extrema(self, dist):
        ext = torch.empty(2, dtype=_f64, device=dist.device)
        _lib.call("ubpl_dist_extrema", dist.data_ptr(), dist.numel(), ext.data_ptr(), _stream())
        return ext

    def reliability(self, dist, legal, ext, reliableDistMin):
        n = dist.numel()
        rel = torch.empty(n, dtype=_f64, device=dist.device)
        keys = torch.empty(n, dtype=torch.int64, device=dist.device)
        _lib.call("ubpl_reliability", dist.data_ptr(), legal.data_ptr(), n, ext.data_ptr(), float(reliableDistMin),
                  rel.data_ptr(), keys.data_ptr(), _stream())
        return rel, keys

    def state(self, k, device):
        return (torch.zeros(1, dtype=torch.int64, device=device), torch.full((1,), k, dtype=torch.int64, device=device),
                torch.empty(65536, dtype=torch.int32, device=device))

    def histogram(self, keys, prefix, shift, hist):
        _lib.call("ubpl_key_histogram", keys.data_ptr(), keys.numel(), prefix.data_ptr(), shift, hist.data_ptr(), 1, _stream())

    def descend(self, hist, shift, prefix, k_rem):
        _lib.call("ubpl_select_descend", hist.data_ptr(), shift, prefix.data_ptr(), k_rem.data_ptr(), 0, _stream())

    def apply(self, rel, J, prefix, reliableThr):
        n = rel.numel()
        dev = rel.device
        enable = torch.empty(n, dtype=torch.uint8, device=dev)
        gate = torch.empty(n, dtype=_f32, device=dev)
        counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
        thr = torch.empty(1, dtype=_f64, device=dev)
        _lib.call("ubpl_select_apply", rel.data_ptr(), n, J, prefix.data_ptr(), float(reliableThr), enable.data_ptr(),
                  gate.data_ptr(), counts.data_ptr(), thr.data_ptr(), _stream())
        return enable, gate, counts, thr


def pair_distance(c1, c2):
    """python-float distance between two coordinate sets [n,2] (utils/process.py:53-54) -> float64 [n]."""
    _need_cuda(c1, c2)
    c1 = c1.to(_f64).contiguous()
    c2 = c2.to(_f64).contiguous()
    n = c1.shape[0]
    out = torch.empty(n, dtype=_f64, device=c1.device)
    _lib.call("ubpl_pair_distance", c1.data_ptr(), c2.data_ptr(), n, out.data_ptr(), _stream())
    return out


def select_quantile_fused(dist, legal, J, k_rank, reliableThr, reliableDistMin, gate=None, p2p=False):
    """ubpl_select_quantile_fused: the whole quantile selection of this rank's items in ONE launch (single CTA).
    p2p=False: the items are the whole population.  p2p=True: the keys of all ranks are exchanged through the
    peer-memory buffer of dist.init_p2p (a collective: every rank calls it the same number of times).
    `gate` = (kps [n,2] f32, S, img_h, img_w, stride, sigma, loss_weight) folds gate_prepare into the launch;
    the result then carries the final gate, `count` and `grad_scale`."""
    import os
    _need_cuda(dist, legal)
    dist = dist.reshape(-1).to(_f64).contiguous()
    legal = legal.reshape(-1)
    if legal.dtype in (torch.uint8, torch.bool):
        legal = legal.to(torch.uint8).contiguous()
        lf, lu = None, legal
    else:
        legal = legal.to(_f64).contiguous()
        lf, lu = legal, None
    n = dist.numel()
    dev = dist.device
    rel = torch.empty(n, dtype=_f64, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    enable = torch.empty(n, dtype=torch.uint8, device=dev)
    g32 = torch.empty(n, dtype=_f32, device=dev)
    counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
    thr = torch.empty(1, dtype=_f64, device=dev)
    ext = torch.empty(2, dtype=_f64, device=dev)
    kps = grad_scale = count = None
    S, img_h, img_w, stride, sigma, lw = 1, 0, 0, 1.0, 1.0, 1.0
    if gate is not None:
        kps, S, img_h, img_w, stride, sigma, lw = gate
        _need_cuda(kps)
        kps = kps.reshape(-1, 2).to(_f32).contiguous()
        grad_scale = torch.empty(1, dtype=_f32, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("ubpl_select_quantile_fused", dist.data_ptr(), _p(lf), _p(lu), n, J, int(k_rank), float(reliableThr),
              float(reliableDistMin), rel.data_ptr(), _p(keys), enable.data_ptr(), g32.data_ptr(), counts.data_ptr(),
              thr.data_ptr(), ext.data_ptr(), _p(kps), int(img_h), int(img_w), float(stride), float(sigma), int(S),
              float(lw), _p(grad_scale), _p(count), 1 if p2p else 0, _stream())
    return dict(reliability=rel, enable=enable, gate=g32, counts=counts, thr=thr, ext=ext, grad_scale=grad_scale,
                count=count, gate_fused=gate is not None)


def select_quantile_emul(dist, legal, J, reliableThr, reliablePCT, reliableDistMin, n_per_rank=None):
    """The multi-GPU selector with its R ranks emulated on ONE GPU (ubpl_select_quantile_emul: R blocks of one
    cooperative launch exchange their histograms through R local exchange buffers exactly as R GPUs do over NVLink).
    dist / legal [R, n]; n_per_rank (optional list) = items rank r really owns (ragged shards).  Returns per-rank
    enable [R, n] (rows valid up to n_per_rank[r]), counts [R, J+1], thr [R], reliability [R, n]."""
    import ctypes
    _need_cuda(dist, legal)
    R, n = dist.shape
    dist = dist.to(_f64).contiguous()
    legal = legal.to(torch.uint8).contiguous()
    dev = dist.device
    npr = [n] * R if n_per_rank is None else [int(x) for x in n_per_rank]
    n_total = sum(npr)
    cap = max(n, 1)
    stride = (int(_lib.lib().ubpl_p2p_buffer_bytes(R, cap)) + 255) // 256 * 256
    xbuf = torch.zeros(R * stride, dtype=torch.uint8, device=dev)
    rel = torch.zeros(R, n, dtype=_f64, device=dev)
    keys = torch.empty(R, n, dtype=torch.int64, device=dev)
    enable = torch.zeros(R, n, dtype=torch.uint8, device=dev)
    g32 = torch.zeros(R, n, dtype=_f32, device=dev)
    counts = torch.zeros(R, J + 1, dtype=torch.int32, device=dev)
    thr = torch.zeros(R, dtype=_f64, device=dev)
    arr = (ctypes.c_int64 * R)(*npr)
    _lib.call("ubpl_select_quantile_emul", dist.data_ptr(), legal.data_ptr(), R, ctypes.cast(arr, ctypes.c_void_p), n, J,
              int((n_total - 1) * reliablePCT), float(reliableThr), float(reliableDistMin), rel.data_ptr(), keys.data_ptr(),
              enable.data_ptr(), g32.data_ptr(), counts.data_ptr(), thr.data_ptr(), xbuf.data_ptr(), stride, cap, _stream())
    status = xbuf.view(R, stride)[:, 68:72].contiguous().view(torch.int32).reshape(R)
    return dict(reliability=rel, enable=enable, gate=g32, counts=counts, thr=thr, status=status)


def select_quantile(dist, legal, J, reliableThr, reliablePCT, reliableDistMin, group=None, n_total=None, backend=None,
                    gate=None):
    """BusinessUtils.filter_pseudo2 (utils/business.py:173-217) on device: min/max normalise,
    reliability = 1 - unc, exact k-th order statistic (k = int((n-1)*pct) from the top) by a
    4-pass 16-bit radix select, enable = reliability > max(reliableThr, kth).

    With a torch.distributed `group` the extrema and the four histograms are all-reduced (NCCL),
    so every rank derives the same global threshold; each rank passes its own shard of items and
    all shards are assumed equal in size unless n_total is given.  `backend` exists for the
    world_size-2 gloo test of this control flow (tests/test_dist_gloo.py); the product always
    uses the CUDA kernels."""
    import os
    be = backend if backend is not None else _CudaSelectBackend()
    n = dist.numel()
    world = 1
    if group is not None:
        import torch.distributed as td
        world = td.get_world_size(group)
    if n_total is None:
        n_total = n * world
    legacy = os.environ.get("UBPL_SELECT", "") == "legacy"
    if backend is None and not legacy and n >= 1:
        from . import dist as _dist
        if world == 1:
            return select_quantile_fused(dist, legal, J, int((n - 1) * reliablePCT), reliableThr, reliableDistMin, gate=gate)
        if _dist.p2p_ready(group):
            # all ranks' keys are exchanged over NVLink peer memory inside the one kernel (no NCCL call)
            return select_quantile_fused(dist, legal, J, int((n_total - 1) * reliablePCT), reliableThr, reliableDistMin,
                                         gate=gate, p2p=True)
    dist, legal = be.prepare(dist, legal)
    if world == 1 and backend is None and 1 <= n <= (1 << 20):
        # single GPU (legacy one-launch kernel): extrema, reliability, radix select and masks
        dev = dist.device
        k = int((n - 1) * reliablePCT)
        rel = torch.empty(n, dtype=_f64, device=dev)
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        enable = torch.empty(n, dtype=torch.uint8, device=dev)
        gate = torch.empty(n, dtype=_f32, device=dev)
        counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
        thr = torch.empty(1, dtype=_f64, device=dev)
        ext = torch.empty(2, dtype=_f64, device=dev)
        _lib.call("ubpl_select_quantile_local", dist.data_ptr(), legal.data_ptr(), n, J, k, float(reliableThr),
                  float(reliableDistMin), rel.data_ptr(), keys.data_ptr(), enable.data_ptr(), gate.data_ptr(),
                  counts.data_ptr(), thr.data_ptr(), ext.data_ptr(), _stream())
        return dict(reliability=rel, enable=enable, gate=gate, counts=counts, thr=thr, ext=ext)
    if world > 1 and backend is None and int(_lib.lib().ubpl_nccl_ranks()) == world:
        # the library's own NCCL communicator (dist.init_nccl): kernels and all-reduces enqueued by ONE call
        return select_quantile_nccl(dist, legal, J, int((n_total - 1) * reliablePCT), reliableThr, reliableDistMin)
    ext = be.extrema(dist)
    if world > 1:
        ext[1:2].neg_()                                       # one MAX all-reduce of (dist_max, -dist_min)
        td.all_reduce(ext, op=td.ReduceOp.MAX, group=group)
        ext[1:2].neg_()
    rel, keys = be.reliability(dist, legal, ext, reliableDistMin)
    if n_total < 1:
        raise IndexError("list index out of range")          # scores[int(-1*pct)] on an empty list
    k = int((n_total - 1) * reliablePCT)                     # utils/business.py:45
    prefix, k_rem, hist = be.state(k, dist.device)
    for shift in (48, 32, 16, 0):
        be.histogram(keys, prefix, shift, hist)
        if world > 1:
            td.all_reduce(hist, op=td.ReduceOp.SUM, group=group)
        be.descend(hist, shift, prefix, k_rem)
    enable, gate, counts, thr = be.apply(rel, J, prefix, reliableThr)
    return dict(reliability=rel, enable=enable, gate=gate, counts=counts, thr=thr, ext=ext)


def select_quantile_nccl(dist, legal, J, k_rank, reliableThr, reliableDistMin):
    """ubpl_select_quantile_dist: this rank's items against the global k-th order statistic (k_rank counted
    over all ranks from the largest reliability), extrema and histograms all-reduced with NCCL."""
    dist = dist.reshape(-1).to(_f64).contiguous()
    legal = legal.reshape(-1).to(_f64).contiguous()
    n = dist.numel()
    dev = dist.device
    rel = torch.empty(n, dtype=_f64, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    enable = torch.empty(n, dtype=torch.uint8, device=dev)
    gate = torch.empty(n, dtype=_f32, device=dev)
    counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
    thr = torch.empty(1, dtype=_f64, device=dev)
    ws = torch.empty(4 + 65536 + 4, dtype=torch.int32, device=dev)      # UBPL_SELECT_WS_BYTES, 8-byte aligned
    _lib.call("ubpl_select_quantile_dist", dist.data_ptr(), legal.data_ptr(), n, J, int(k_rank), float(reliableThr),
              float(reliableDistMin), rel.data_ptr(), keys.data_ptr(), enable.data_ptr(), gate.data_ptr(),
              counts.data_ptr(), thr.data_ptr(), ws.data_ptr(), _stream())
    return dict(reliability=rel, enable=enable, gate=gate, counts=counts, thr=thr, ext=ws[:4].view(_f64))


def mix_dists(p1, s1, a1, p2, s2, a2, gt=None, pck_ref=(0, 1), pck_thr=0.2):
    """The per-key-point quantities of BusinessUtils.pseudo_cal_unc (utils/business.py:220-234,302-326) for two
    teachers: p [B,J,2], s [B,J] (or None), a [B,J,A,2] (the A augmented views of every key point), gt [B,J,>=2].
    Returns float64 [B,J] tensors err1/2, score1/2, int1/2, ext, aext, caug1/2 [B,J,2] and int32 acc1/2."""
    _need_cuda(p1, p2, a1, a2, s1, s2, gt)
    p1, p2 = p1.to(_f32).contiguous(), p2.to(_f32).contiguous()
    a1, a2 = a1.to(_f32).contiguous(), a2.to(_f32).contiguous()
    B, J, A, _ = a1.shape
    dev = a1.device
    s1 = None if s1 is None else s1.to(_f32).contiguous()
    s2 = None if s2 is None else s2.to(_f32).contiguous()
    gt = None if gt is None else gt.to(_f32).contiguous()
    o = {k: torch.empty(B, J, dtype=_f64, device=dev) for k in ("int1", "int2", "ext", "aext")}
    o["caug1"] = torch.empty(B, J, 2, dtype=_f64, device=dev)
    o["caug2"] = torch.empty(B, J, 2, dtype=_f64, device=dev)
    for m in ("1", "2"):
        o["err" + m] = torch.empty(B, J, dtype=_f64, device=dev) if gt is not None else None
        o["acc" + m] = torch.empty(B, J, dtype=torch.int32, device=dev) if gt is not None else None
        o["score" + m] = torch.empty(B, J, dtype=_f64, device=dev) if (s1 if m == "1" else s2) is not None else None
    _lib.call("ubpl_mix_dists", _p(gt), gt.shape[-1] if gt is not None else 0, int(pck_ref[0]), int(pck_ref[1]), float(pck_thr),
              p1.data_ptr(), p2.data_ptr(), _p(s1), _p(s2), a1.data_ptr(), a2.data_ptr(), B, J, A,
              _p(o["err1"]), _p(o["err2"]), _p(o["acc1"]), _p(o["acc2"]), _p(o["score1"]), _p(o["score2"]),
              o["caug1"].data_ptr(), o["caug2"].data_ptr(), o["int1"].data_ptr(), o["int2"].data_ptr(),
              o["ext"].data_ptr(), o["aext"].data_ptr(), _stream())
    return o


class MixUncState:
    """Device form of args.mdsN_lma_cache (utils/business.py:348-355,378-393) for n key points of ONE teacher:
    the last three (intDist, extDist, aExtDist) of every key point, updated in place by mix_unc."""

    def __init__(self, n, device):
        self.n = int(n)
        self.hist = torch.zeros(3, self.n, 3, dtype=_f64, device=device)
        self.len = torch.zeros(self.n, dtype=torch.int32, device=device)


def mix_unc(intDist, extDist, aExtDist, J, distThrMax, state, score=None, score_thr=None):
    """LMA (0.5/0.3/0.2) of the three distances -> mixDist -> unc (utils/business.py:320-346,395-405) and the
    fixed rule of pseudo_filter_mixUnc (:237-261); score/score_thr add the gate of pseudo_filter_mixUnc2."""
    _need_cuda(intDist, extDist, aExtDist, score, score_thr)
    i, e, a = (t.reshape(-1).to(_f64).contiguous() for t in (intDist, extDist, aExtDist))
    n = i.numel()
    if n != state.n:
        raise ValueError("state holds %d key points, got %d" % (state.n, n))
    dev = i.device
    score = None if score is None else score.reshape(-1).to(_f64).contiguous()
    score_thr = None if score_thr is None else score_thr.reshape(-1).to(_f64).contiguous()
    lma = torch.empty(3, n, dtype=_f64, device=dev)
    mix = torch.empty(n, dtype=_f64, device=dev)
    unc = torch.empty(n, dtype=_f64, device=dev)
    enable = torch.empty(n, dtype=torch.uint8, device=dev)
    gate = torch.empty(n, dtype=_f32, device=dev)
    counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
    _lib.call("ubpl_mix_unc", i.data_ptr(), e.data_ptr(), a.data_ptr(), n, int(J), float(distThrMax), state.hist.data_ptr(),
              state.len.data_ptr(), _p(score), _p(score_thr), lma.data_ptr(), mix.data_ptr(), unc.data_ptr(),
              enable.data_ptr(), gate.data_ptr(), counts.data_ptr(), _stream())
    return dict(intDist_lma=lma[0], extDist_lma=lma[1], aExtDist_lma=lma[2], mixDist=mix, unc=unc, enable=enable, gate=gate,
                counts=counts)


def select_fixed(dist, legal, J, distThrMax):
    """enable = legal and 1-exp(-dist/5) <= 1-exp(-3*distThrMax/5)  (utils/business.py:237-261)."""
    _need_cuda(dist, legal)
    dist = dist.reshape(-1).to(_f64).contiguous()
    legal = None if legal is None else legal.reshape(-1).to(_f64).contiguous()
    n = dist.numel()
    dev = dist.device
    enable = torch.empty(n, dtype=torch.uint8, device=dev)
    gate = torch.empty(n, dtype=_f32, device=dev)
    counts = torch.empty(J + 1, dtype=torch.int32, device=dev)
    unc = torch.empty(n, dtype=_f64, device=dev)
    _lib.call("ubpl_select_fixed", dist.data_ptr(), _p(legal), n, J, float(distThrMax), enable.data_ptr(),
              gate.data_ptr(), counts.data_ptr(), unc.data_ptr(), _stream())
    return dict(enable=enable, gate=gate, counts=counts, unc=unc)


def k2_view_fixed(preds, distThrMax, S, img_h, img_w, stride, sigma=3.0):
    """One-launch K2 for the mean-teacher fixed-threshold path: dispersion + fixed rule + visibility gate +
    open-gate count.  preds [K,B,J,2].  `count` (= S * #(gate > 0)) feeds render_mse(count_in=...)."""
    _need_cuda(preds)
    preds = preds.to(_f32).contiguous()
    K, B, J, _ = preds.shape
    dev = preds.device
    mean = torch.empty(B, J, 2, dtype=_f32, device=dev)
    dist = torch.empty(B, J, dtype=_f64, device=dev)
    legal = torch.empty(B, J, dtype=torch.uint8, device=dev)
    enable = torch.empty(B, J, dtype=torch.uint8, device=dev)
    gate = torch.empty(B, J, dtype=_f32, device=dev)
    counts = torch.empty(J + 2, dtype=torch.int32, device=dev)
    count = counts[J + 1:]
    _lib.call("ubpl_k2_view_fixed", preds.data_ptr(), K, B, J, float(distThrMax), int(img_h), int(img_w), float(stride),
              float(sigma), int(S), mean.data_ptr(), dist.data_ptr(), legal.data_ptr(), enable.data_ptr(),
              gate.data_ptr(), count.data_ptr(), counts.data_ptr(), _stream())
    return dict(mean=mean, dist=dist, legal=legal, enable=enable, gate=gate, count=count, counts=counts[:J + 1])


# -------------------------------------------------------------------------------------------------
# K3
# -------------------------------------------------------------------------------------------------
def render_targets(kps, H, W, img_h, img_w, stride=None, sigma=3.0):
    """ProcessUtils.kps_heatmap (utils/process.py:253-278) for N key points: kps [N,3] ->
    (heatmap [N,H,W], kps_out [N,3] with weight *= visibility)."""
    _need_cuda(kps)
    kps = kps.to(_f32).contiguous()
    N = kps.shape[0]
    if stride is None:
        stride = img_w / W
    hm = torch.empty(N, H, W, dtype=_f32, device=kps.device)
    kout = torch.empty(N, 3, dtype=_f32, device=kps.device)
    _lib.call("ubpl_render_targets", kps.data_ptr(), N, H, W, int(img_h), int(img_w), float(stride), float(sigma),
              hm.data_ptr(), kout.data_ptr(), _stream())
    return hm, kout


_sum_ws = {}
_SUM_WS_WORDS = (16 + 24 * 4096) // 4 + 12          # UBPL_RENDER_SUM_WS_BYTES, padded to a multiple of 64 B


_GRAPH_WS_POOL = 16


def _sum_workspace(dev):
    """A zero-initialised workspace for kernels that elect their last CTA (the kernel returns its ticket word to
    zero, so a buffer can serve any number of launches that do not overlap in time).  Eager launches use ONE buffer
    per (device, stream): launches on a stream are ordered, so they never share a live ticket.  A launch that is being
    captured into a CUDA graph takes a buffer of its own from a pool zeroed beforehand -- the graph keeps replaying
    with it while eager calls and other graphs never touch it (when the pool is empty the buffer is zeroed inside the
    capture, one extra memset node)."""
    didx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (dev.type, didx)
    st = _sum_ws.get(key)
    if st is None:
        st = _sum_ws[key] = {"streams": {}, "pool": []}
    if torch.cuda.is_current_stream_capturing():
        if st["pool"]:
            return st["pool"].pop()
        return torch.zeros(_SUM_WS_WORDS, dtype=torch.int32, device=dev)
    if not st["pool"]:
        st["pool"] = list(torch.zeros(_GRAPH_WS_POOL, _SUM_WS_WORDS, dtype=torch.int32, device=dev).unbind(0))
    sid = torch.cuda.current_stream(didx).cuda_stream
    ws = st["streams"].get(sid)
    if ws is None:
        ws = st["streams"][sid] = torch.zeros(_SUM_WS_WORDS, dtype=torch.int32, device=dev)
    return ws


def render_mse(kps, gate, sample_w, pred, img_h, img_w, stride=None, sigma=3.0, grad_scale=None,
               want_grad=True, want_target=True, count_in=None, loss_weight=1.0, want_summary=False):
    """Fused Gaussian render + JointMSELoss forward + gradient (K3b).  kps [B,J,2] image space,
    gate [B,J] or None, sample_w [B] / [B,1] or None, pred [B,S,J,H,W].  Returns dict(per_loss
    [B,S,J], gate_out [B,J], grad, target[, summary]); want_summary adds the float64[4] reduction of
    loss_finalize(per_loss, None, gate_out) computed by the same launch."""
    _need_cuda(kps, gate, sample_w, pred, grad_scale)
    pred = _inner_contig(pred)
    B, S, J, H, W = pred.shape
    dev = pred.device
    kps = kps.reshape(B, J, 2).to(_f32).contiguous()
    gate = None if gate is None else gate.reshape(B, J).to(_f32).contiguous()
    sample_w = None if sample_w is None else sample_w.reshape(B).to(_f32).contiguous()
    if stride is None:
        stride = img_w / W
    grad = torch.empty(B, S, J, H, W, dtype=_f32, device=dev) if want_grad else None
    target = torch.empty(B, J, H, W, dtype=_f32, device=dev) if want_target else None
    gate_out = torch.empty(B, J, dtype=_f32, device=dev)
    per_loss = torch.empty(B, S, J, dtype=_f32, device=dev)
    gs_out = torch.empty(1, dtype=_f32, device=dev) if count_in is not None else None
    gs = (0, 0, 0) if grad is None else (grad.stride(0), grad.stride(1), grad.stride(2))
    args = (kps.data_ptr(), _p(gate), _p(sample_w), pred.data_ptr(), pred.stride(0),
            pred.stride(1), pred.stride(2), _p(grad), gs[0], gs[1], gs[2], _p(target), B, S, J, H, W, int(img_h),
            int(img_w), float(stride), float(sigma), _p(grad_scale), _p(count_in), float(loss_weight), _p(gs_out),
            gate_out.data_ptr(), per_loss.data_ptr())
    summary = None
    if want_summary:
        summary = torch.empty(4, dtype=_f64, device=dev)
        _lib.call("ubpl_render_mse_sum", *args, summary.data_ptr(), _sum_workspace(dev).data_ptr(), _stream())
    else:
        _lib.call("ubpl_render_mse", *args, _stream())
    return dict(per_loss=per_loss, gate_out=gate_out, grad=grad, target=target, grad_scale=gs_out, summary=summary)


def dense_mse(pred, tgt, coef=None, mask_mode=0, thr=0.0, grad_scale=None, want_grad=True, want_scores=False):
    """Dense-target masked joint-MSE forward + gradient (K3a/K3c).  pred [B,S,J,H,W];
    tgt [M,B,S,J,H,W] (per-stack targets) or [M,B,J,H,W] (shared by the stacks); coef [B,J]."""
    _need_cuda(pred, tgt, coef, grad_scale)
    pred = _inner_contig(pred)
    tgt = _inner_contig(tgt)
    B, S, J, H, W = pred.shape
    dev = pred.device
    M = tgt.shape[0]
    if tgt.dim() == 6:
        tM, tB, tS, tJ = tgt.stride(0), tgt.stride(1), tgt.stride(2), tgt.stride(3)
    else:
        tM, tB, tS, tJ = tgt.stride(0), tgt.stride(1), 0, tgt.stride(2)
    coef = None if coef is None else coef.reshape(B, J).to(_f32).contiguous()
    grad = torch.empty(B, S, J, H, W, dtype=_f32, device=dev) if want_grad else None
    per_loss = torch.empty(B, S, J, dtype=_f32, device=dev)
    mask = torch.empty(B, S, J, dtype=_f32, device=dev)
    vp = torch.empty(B, S, J, dtype=_f32, device=dev) if (want_scores or mask_mode == 1) else None
    vt = torch.empty(B, S, J, dtype=_f32, device=dev) if (want_scores or mask_mode != 0) else None
    gs = (0, 0, 0) if grad is None else (grad.stride(0), grad.stride(1), grad.stride(2))
    _lib.call("ubpl_dense_mse", pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), tgt.data_ptr(), M,
              tM, tB, tS, tJ, _p(coef), int(mask_mode), float(thr), _p(grad), gs[0], gs[1], gs[2], B, S, J, H, W,
              _p(grad_scale), per_loss.data_ptr(), mask.data_ptr(), _p(vp), _p(vt), _stream())
    return dict(per_loss=per_loss, mask=mask, vmax_p=vp, vmax_t=vt, grad=grad)


def loss_finalize(per_loss, mask=None, gate=None):
    """float64[4] on device: sum(per_loss*mask), #(per_loss>0), #(mask>0), #(gate>0) (B*J if None)."""
    _need_cuda(per_loss, mask, gate)
    B, S, J = per_loss.shape
    out = torch.empty(4, dtype=_f64, device=per_loss.device)
    _lib.call("ubpl_loss_finalize", per_loss.data_ptr(), _p(mask), _p(gate), B, S, J, out.data_ptr(), _stream())
    return out


def gate_prepare(kps, gate_in, S, img_h, img_w, stride, sigma=3.0, loss_weight=1.0):
    """gate_out = gate_in*visibility, count = S*#(gate_out>0), grad_scale = loss_weight/count, all on
    device (utils/process.py:262-268, utils/losses.py:29, projects/MT_UBPL.py:266)."""
    _need_cuda(kps, gate_in)
    kps = kps.reshape(-1, 2).to(_f32).contiguous()
    n = kps.shape[0]
    gate_in = None if gate_in is None else gate_in.reshape(-1).to(_f32).contiguous()
    dev = kps.device
    gate_out = torch.empty(n, dtype=_f32, device=dev)
    grad_scale = torch.empty(1, dtype=_f32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("ubpl_gate_prepare", kps.data_ptr(), _p(gate_in), n, int(img_h), int(img_w), float(stride), float(sigma),
              int(S), float(loss_weight), gate_out.data_ptr(), grad_scale.data_ptr(), count.data_ptr(), _stream())
    return gate_out, grad_scale, count


def scale(x, scale, out=None):
    """out = x * scale with `scale` a float32 device scalar (no host sync); out=x scales in place."""
    _need_cuda(x, scale)
    assert x.is_contiguous() and x.dtype == _f32
    if out is None:
        out = torch.empty_like(x)
    _lib.call("ubpl_scale", out.data_ptr(), x.data_ptr(), x.numel(), scale.data_ptr(), _stream())
    return out


def scale_inplace(x, s):
    return scale(x, s, out=x)


# -------------------------------------------------------------------------------------------------
# N2 / N3
# -------------------------------------------------------------------------------------------------
def view_kps(kps, mats, flips, img_w):
    """Canonical key points [B,J,3] -> every augmented view's frame [V,B,J,3] (utils/process.py:239-242,
    utils/augment.py:151-156, utils/udaap/transforms.py:151-158).  mats [V,B,3,3] or [V,B,2,3] float64 from
    augment.AugmentUtils.view_matrix (get_transform of the view), flips [V,B] bool/uint8 or None."""
    _need_cuda(kps, mats, flips)
    kps = kps.to(_f32).contiguous()
    B, J, _ = kps.shape
    V = mats.shape[0]
    mats = mats.to(_f64)[..., :2, :].contiguous()
    flips = None if flips is None else flips.reshape(V, B).to(torch.uint8).contiguous()
    out = torch.empty(V, B, J, 3, dtype=_f32, device=kps.device)
    _lib.call("ubpl_view_kps", kps.data_ptr(), mats.data_ptr(), _p(flips), float(img_w), V, B, J, out.data_ptr(), _stream())
    return out


def acc_pck(preds, gts, pck_ref, pck_thr, want_dists=False):
    """EvaluationUtils.acc_pck (utils/evaluation.py:92-139): preds [bs,k,>=2], gts [bs,k,>=2] ->
    (errs [k+1], accs [k+1]) float32 on the device (+ dists, dists_ref [k,bs] when asked)."""
    _need_cuda(preds, gts)
    preds = preds.to(_f32).contiguous()
    gts = gts.to(_f32).contiguous()
    bs, k = preds.shape[:2]
    dev = preds.device
    errs = torch.empty(k + 1, dtype=_f32, device=dev)
    accs = torch.empty(k + 1, dtype=_f32, device=dev)
    dists = torch.empty(k, bs, dtype=_f32, device=dev) if want_dists else None
    dref = torch.empty(k, bs, dtype=_f32, device=dev) if want_dists else None
    _lib.call("ubpl_acc_pck", preds.data_ptr(), preds.shape[-1], gts.data_ptr(), gts.shape[-1], bs, k, int(pck_ref[0]),
              int(pck_ref[1]), float(pck_thr), errs.data_ptr(), accs.data_ptr(), _p(dists), _p(dref), _stream())
    return (errs, accs, dists, dref) if want_dists else (errs, accs)


def features_cov(inp1, inp2, want_grad=True):
    """ProcessUtils.features_cov (utils/process.py:19-31) forward + gradient: inp [bs,n,c,h,w] ->
    dict(value 0-d float32, rows, cov [bs,n,c], grad1, grad2)."""
    _need_cuda(inp1, inp2)
    a = inp1.to(_f32).contiguous()
    b = inp2.to(_f32).contiguous()
    bs, n, c, h, w = a.shape
    rows, L = bs * n * c, h * w
    dev = a.device
    cov = torch.empty(bs, n, c, dtype=_f32, device=dev)
    value = torch.empty(1, dtype=_f32, device=dev)
    g1 = torch.empty_like(a) if want_grad else None
    g2 = torch.empty_like(b) if want_grad else None
    _lib.call("ubpl_features_cov", a.data_ptr(), b.data_ptr(), rows, L, cov.data_ptr(), value.data_ptr(), _p(g1), _p(g2), _stream())
    return dict(value=value[0], rows=rows, cov=cov, grad1=g1, grad2=g2)


# -------------------------------------------------------------------------------------------------
# K4
# -------------------------------------------------------------------------------------------------
def ema_work_items(numels, size):
    """The (tensor, first element) table of an EMA launch: every tensor cut into work items of at most `size` elements,
    in tensor order.  Pure host function (tested on CPU): the items of a tensor are disjoint and cover it exactly."""
    which, start = [], []
    for t, n in enumerate(numels):
        for s in range(0, int(n), int(size)):
            which.append(t)
            start.append(s)
    return which, start


class EmaPlan:
    """Device-side pointer/chunk tables for the one-launch EMA of a (student, teacher) model pair
    (utils/parameters.py:4-8).  Rebuilt automatically if any parameter storage moved."""
    CHUNK = 8192          # elements per work item of the standalone kernel (one CTA sweep)
    PIECE = 1024          # elements per work item when the update rides in K1's launch (one warp, one memory round trip;
                          # kEmaPiece in csrc/warp_decode.cu): its own table, so that a small tensor is ONE claim

    def __init__(self, params, ema_params):
        self.params = list(params)
        self.ema_params = list(ema_params)
        if len(self.params) != len(self.ema_params):
            raise ValueError("student and teacher have different parameter counts")
        self._key = None
        self._build()

    def _ptr_key(self):
        return tuple(p.data_ptr() for p in self.params) + tuple(e.data_ptr() for e in self.ema_params)

    def _build(self):
        dev = self.ema_params[0].device
        _need_cuda(*self.params, *self.ema_params)
        for p, e in zip(self.params, self.ema_params):
            if p.dtype != _f32 or e.dtype != _f32 or not p.is_contiguous() or not e.is_contiguous():
                raise _lib.UbplError("EMA expects contiguous float32 parameters")
            if p.numel() != e.numel():
                raise ValueError("parameter shape mismatch between student and teacher")
        numels = [p.numel() for p in self.params]
        ct, cs = ema_work_items(numels, self.CHUNK)
        self.n_chunks = len(ct)
        self.n_elems = sum(numels)
        self.ema_ptrs = torch.tensor([e.data_ptr() for e in self.ema_params], dtype=torch.int64, device=dev)
        self.param_ptrs = torch.tensor([p.data_ptr() for p in self.params], dtype=torch.int64, device=dev)
        self.numels = torch.tensor(numels, dtype=torch.int64, device=dev)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self.chunk_start = torch.tensor(cs, dtype=torch.int64, device=dev)
        pt, ps = ema_work_items(numels, self.PIECE)
        self.n_pieces = len(pt)
        self.piece_tensor = torch.tensor(pt, dtype=torch.int32, device=dev)
        self.piece_start = torch.tensor(ps, dtype=torch.int64, device=dev)
        self._key = self._ptr_key()

    def set_alpha(self, alpha):
        """Writes {alpha, 1 - alpha} (float32) into the plan's device buffer; launches made with
        step(alpha, from_device=True) -- in particular launches captured into a CUDA graph -- read it at run time."""
        import numpy as np
        a = np.float32(alpha)
        vals = torch.tensor([float(a), float(np.float32(1 - alpha))], dtype=_f32)
        if getattr(self, "alpha_buf", None) is None:
            self.alpha_buf = torch.empty(2, dtype=_f32, device=self.ema_params[0].device)
        self.alpha_buf.copy_(vals, non_blocking=False)
        self.alpha = float(alpha)

    def launch_args(self, alpha, from_device=False, pieces=False):
        """The EMA arguments shared by ubpl_ema_multi_tensor and ubpl_warp_decode_k2_ema (tables, chunking, alpha);
        pieces=True: the finer table for the launch that rides in K1."""
        if self._ptr_key() != self._key:
            self._build()
        import numpy as np
        a = float(np.float32(alpha))
        oma = float(np.float32(1 - alpha))
        adev = None
        if from_device:
            if getattr(self, "alpha_buf", None) is None:
                self.set_alpha(alpha)
            adev = self.alpha_buf.data_ptr()
        if pieces:
            return (self.ema_ptrs.data_ptr(), self.param_ptrs.data_ptr(), self.numels.data_ptr(), self.piece_tensor.data_ptr(),
                    self.piece_start.data_ptr(), self.n_pieces, self.PIECE, a, oma, adev)
        return (self.ema_ptrs.data_ptr(), self.param_ptrs.data_ptr(), self.numels.data_ptr(), self.chunk_tensor.data_ptr(),
                self.chunk_start.data_ptr(), self.n_chunks, self.CHUNK, a, oma, adev)

    def step(self, alpha, from_device=False):
        """ema <- ema*alpha + (1-alpha)*param over every tensor, one launch.  from_device=True: alpha is read from the
        device buffer written by set_alpha (the scalar argument is then only the fallback when no buffer exists)."""
        _lib.call("ubpl_ema_multi_tensor", *self.launch_args(alpha, from_device), _stream())


def ema_flat(ema, param, alpha):
    _need_cuda(ema, param)
    assert ema.is_contiguous() and param.is_contiguous() and ema.dtype == _f32 and param.dtype == _f32
    import numpy as np
    _lib.call("ubpl_ema_flat", ema.data_ptr(), param.data_ptr(), ema.numel(), float(np.float32(alpha)),
              float(np.float32(1 - alpha)), _stream())
    return ema
