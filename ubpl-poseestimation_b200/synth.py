"""Seeded synthetic inputs for the pseudo-label hot path.

Shapes follow the reference's tensor contract: teacher stacks
outs_ema[M,K,B,J,H,W] (last hourglass stack only), student outs[B,S,J,H,W], per-view
meta['warpmat'][K,B,2,3] / meta['isflip'][K,B] as produced by datasets/dataset_mds.py:117,197-200
(theta = (1/s)*[[cos a, sin a, 0], [-sin a, cos a, 0]] for affine_getWarpmat(-a, 1/s),
utils/augment.py:159-164), centre [128,128] (utils/process.py:218-221).

Everything is generated with torch ops on the requested device so the B200 benchmark can build
its batches in HBM; the tests call it with device='cpu' and small sizes.
"""
import math
import torch


def make_warpmats(K, B, gen, device, sf=0.25, rf=30.0, base_scale=1.28):
    """angle ~ clamp(N(0,rf), +-rf) degrees, s = base*clamp(1+N(0,sf), 1-sf, 1+sf)
    (utils/augment.py:20-21); flip ~ Bernoulli(0.5) (utils/augment.py:185-192)."""
    ang = (torch.randn(K, B, generator=gen, device=device) * rf).clamp(-rf, rf) * (math.pi / 180.0)
    s = base_scale * (1.0 + torch.randn(K, B, generator=gen, device=device) * sf).clamp(1 - sf, 1 + sf)
    theta = torch.zeros(K, B, 2, 3, device=device, dtype=torch.float32)
    theta[..., 0, 0] = torch.cos(ang) / s
    theta[..., 0, 1] = torch.sin(ang) / s
    theta[..., 1, 0] = -torch.sin(ang) / s
    theta[..., 1, 1] = torch.cos(ang) / s
    flip = torch.rand(K, B, generator=gen, device=device) < 0.5
    return theta, flip, s


def _blobs(cx, cy, amp, H, W, device):
    """amp*exp(-((x-cx)^2+(y-cy)^2)/18) on an HxW grid; cx/cy/amp broadcast over leading dims."""
    ys = torch.arange(H, device=device, dtype=torch.float32).view(H, 1)
    xs = torch.arange(W, device=device, dtype=torch.float32).view(1, W)
    d2 = (xs - cx[..., None, None]) ** 2 + (ys - cy[..., None, None]) ** 2
    return amp[..., None, None] * torch.exp(-d2 / 18.0)


def canonical_to_view(cx, cy, theta, flip, H, W):
    """Source-frame (view) pixel that the back-warp of utils/augment.py:37-47 brings to the
    canonical pixel (cx, cy): un-mirror x, normalise, apply theta, un-normalise."""
    xo = torch.where(flip, (W - 1) - cx, cx)
    xn = -1.0 + 2.0 * xo / (W - 1)
    yn = -1.0 + 2.0 * cy / (H - 1)
    xs = theta[..., 0, 0] * xn + theta[..., 0, 1] * yn + theta[..., 0, 2]
    ys = theta[..., 1, 0] * xn + theta[..., 1, 1] * yn + theta[..., 1, 2]
    return (xs + 1.0) * 0.5 * (W - 1), (ys + 1.0) * 0.5 * (H - 1)


def make_batch(B, K, J, H=64, W=64, M=1, S=2, seed=1388, rank=0, device="cpu",
               neg_frac=0.05, jitter=1.5, noise=0.02, chunk=64, noise_only_frac=0.2):
    """Returns a dict with teacher[M,K,B,J,H,W], student[B,S,J,H,W], theta[K,B,2,3],
    flip[K,B] (bool), center[B,2] (int64), scale[B] (float32), islabeled[B] (bool; the LAST B/2
    rows are labeled, utils/mt/data.py:121-129), base_xy[B,J,2].

    neg_frac of the teacher maps are all-negative (they exercise the max <= 0 mask of
    utils/udaap/evaluation.py:27-29): most of them are weak responses -- the same smooth blob shifted
    below zero, which is what a heat-map head emits for a low-confidence joint -- and a share
    noise_only_frac of them is pure negative noise (no structure at all; the decoder's worst case)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + rank)
    theta, flip, _ = make_warpmats(K, B, gen, device)
    base = 8.0 + torch.rand(B, J, 2, generator=gen, device=device) * (torch.tensor([W - 16.0, H - 16.0], device=device))
    teacher = torch.empty(M, K, B, J, H, W, device=device, dtype=torch.float32)
    student = torch.empty(B, S, J, H, W, device=device, dtype=torch.float32)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        jit = torch.randn(M, K, nb, J, 2, generator=gen, device=device) * jitter
        cx = base[b0:b1, :, 0] + jit[..., 0]
        cy = base[b0:b1, :, 1] + jit[..., 1]
        th = theta[None, :, b0:b1, None]          # [1,K,nb,1,2,3]
        fl = flip[None, :, b0:b1, None]
        qx, qy = canonical_to_view(cx, cy, th, fl, H, W)
        amp = 0.5 + 0.6 * torch.rand(M, K, nb, J, generator=gen, device=device)
        t = _blobs(qx, qy, amp, H, W, device)
        nz = torch.randn(M, K, nb, J, H, W, generator=gen, device=device) * noise
        neg = torch.rand(M, K, nb, J, generator=gen, device=device) < neg_frac
        noise_only = torch.rand(M, K, nb, J, generator=gen, device=device) < noise_only_frac
        weak = t + nz - (amp[..., None, None] + 0.2)                    # max = amp + noise - amp - 0.2 < 0
        negmap = torch.where(noise_only[..., None, None], -(nz.abs() + 1e-3), weak)
        t = torch.where(neg[..., None, None], negmap, t + nz)
        teacher[:, :, b0:b1] = t
        sj = torch.randn(nb, S, J, 2, generator=gen, device=device) * jitter
        samp = 0.5 + 0.6 * torch.rand(nb, S, J, generator=gen, device=device)
        st = _blobs(base[b0:b1, None, :, 0] + sj[..., 0], base[b0:b1, None, :, 1] + sj[..., 1], samp, H, W, device)
        st = st + torch.randn(nb, S, J, H, W, generator=gen, device=device) * noise
        student[b0:b1] = st
    center = torch.full((B, 2), 2 * W, dtype=torch.int64, device=device)   # 128 for 64x64 maps (256 input)
    scale = torch.full((B,), 1.28 * (W / 64.0), dtype=torch.float32, device=device)
    islabeled = torch.zeros(B, dtype=torch.bool, device=device)
    islabeled[B - B // 2:] = True
    return dict(teacher=teacher, student=student, theta=theta, flip=flip, center=center, scale=scale,
                islabeled=islabeled, base_xy=base)
