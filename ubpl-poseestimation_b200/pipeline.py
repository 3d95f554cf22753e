"""The fused pseudo-label chain, device resident, no host sync:

    K1  back-warp + flip + arg-max decode of every teacher view      (augment.py:37-47, process.py:330-336)
    K2  per-joint dispersion -> selection mask                        (evaluation.py:40-58 / business.py:109-217)
    K3  Gaussian target render + masked joint-MSE forward + gradient  (process.py:253-278, losses.py:8-29)

`pseudo_label_step` composes the reference's library functions in the order of SURVEY 3.4:
kps_fromHeatmap_mul on the back-warped maps of every teacher -> (M=1) uncertainty_fromDistance
dispersion, or (M=2) assess_pseudo_unc2 on [predsMean_1, predsMean_2, preds_mean] ->
filter_pseudo2 (global quantile) or the fixed rule of pseudo_filter_mixUnc -> Dataset.update
(coords + enable) -> kps_heatmap -> JointMSELoss(useKPsGate, useSampleWeight) with the
`getSampleWeight_nega` weights (unlabeled rows get pseudoWeight) -> weight * sum / n
(MT_UBPL.py:266) and its gradient."""
from dataclasses import dataclass

import torch

from . import ops


@dataclass
class StepConfig:
    select: str = "fixed"            # "fixed" (business.py:237-261) or "quantile" (business.py:173-217)
    distThrMax: float = 1.0          # fixed rule: unc <= 1-exp(-3*distThrMax/5)
    reliableThr: float = 0.0
    reliablePCT: float = 0.5
    reliableDistMin: float = 1.0
    pseudoWeight: float = 1.0        # projects/tools.py:24-31
    lossWeight: float = 1.0          # args.poseWeight, MT_UBPL.py:266
    stride: float = 4.0              # inpRes / outRes, process.py:255
    sigma: float = 3.0               # kernelSize * sigma, process.py:258
    want_target: bool = True         # materialise the rendered targets (counted in the byte model)
    want_grad: bool = True
    fuse_k2: bool = True             # one-launch K2 for the M=1 fixed-threshold path


def nega_weights(islabeled, pseudoWeight):
    """projects/tools.py:24-31 getSampleWeight_nega: labeled rows 0, unlabeled rows pseudoWeight."""
    return torch.where(islabeled.bool(), torch.zeros((), device=islabeled.device),
                       torch.full((), float(pseudoWeight), device=islabeled.device)).to(torch.float32)


def pseudo_label_step(teacher, student, theta, flip, dec, sample_w, cfg: StepConfig, group=None, stats=None,
                      timer=None):
    """teacher [M,K,B,J,H,W] (last-stack teacher maps of the K augmented views), student
    [B,S,J,H,W], theta [K,B,2,3], flip [K,B], dec [B,4] (ops.decode_coeffs), sample_w [B].
    Returns a dict of DEVICE tensors: summary float64[4] = (loss_sum, #loss>0, #mask>0, #gate>0),
    grad_scale, count, grad, target, gate, kps, enable, dist, decode outputs.  The scalar loss of
    MT_UBPL.py:266 is summary[0] * grad_scale."""
    M, K, B, J, H, W = teacher.shape
    S = student.shape[1]
    stride = cfg.stride
    img_h, img_w = int(H * stride), int(W * stride)
    mark = timer if timer is not None else (lambda name: None)   # bench.py records CUDA events at stage edges
    # ---- K1: every (model, view) map read once ---------------------------------------------------
    mark("k1_0")
    if teacher.stride(0) == K * teacher.stride(1):
        dec_out = ops.warp_decode(teacher.view(M * K, B, J, H, W) if teacher.is_contiguous() else
                                  teacher.as_strided((M * K, B, J, H, W), (teacher.stride(1),) + tuple(teacher.stride()[2:])),
                                  theta.unsqueeze(0).expand(M, K, B, 2, 3).reshape(M * K, B, 2, 3),
                                  flip.unsqueeze(0).expand(M, K, B).reshape(M * K, B), dec, stats=stats, want_idx=True)
        xy = dec_out["xy"].view(M, K, B, J, 2)
        mx = dec_out["max"].view(M, K, B, J)
        idx = dec_out["idx"].view(M, K, B, J)
    else:
        outs = [ops.warp_decode(teacher[m], theta, flip, dec, stats=stats) for m in range(M)]
        xy = torch.stack([o["xy"] for o in outs])
        mx = torch.stack([o["max"] for o in outs])
        idx = torch.stack([o["idx"] for o in outs])
    mark("k1_1")
    # ---- K2: dispersion + selection ----------------------------------------------------------------
    fused_k2 = (cfg.fuse_k2 and M == 1 and cfg.select == "fixed" and B * J <= 65536)
    if fused_k2:
        k2 = ops.k2_view_fixed(xy[0], cfg.distThrMax, S, img_h, img_w, stride, cfg.sigma)
        kps, dist, legal = k2["mean"], k2["dist"], k2["legal"]
        gate, grad_scale, count = k2["gate"], None, k2["count"]
        sel = dict(enable=k2["enable"], counts=k2["counts"])
        extra = {}
    else:
        if M == 1:
            vd = ops.view_dispersion(xy[0], sentinel_illegal=True)
            kps, dist, legal = vd["mean"], vd["dist"], vd["legal"]
            extra = dict(unc32=vd["unc32"], max_bits=vd["max_bits"])
        elif M == 2:
            vd1, vd2 = ops.view_dispersion(xy[0]), ops.view_dispersion(xy[1])
            ad = ops.assess_dual(vd1["mean"], vd2["mean"], None, xy[0], xy[1])
            kps, dist, legal = ad["coord32"], ad["extDist"], ad["legal"]
            extra = dict(assess=ad)
        else:
            raise ValueError("pseudo_label_step supports M = 1 (mean teacher) or M = 2 (dual teachers)")
        if cfg.select == "fixed":
            sel = ops.select_fixed(dist, legal, J, cfg.distThrMax)
        elif cfg.select == "quantile":
            sel = ops.select_quantile(dist, legal, J, cfg.reliableThr, cfg.reliablePCT, cfg.reliableDistMin, group=group)
        else:
            raise ValueError("select must be 'fixed' or 'quantile'")
        # ---- K3: render + masked MSE forward/backward ----------------------------------------------
        gate, grad_scale, count = ops.gate_prepare(kps, sel["gate"], S, img_h, img_w, stride, cfg.sigma, cfg.lossWeight)
    mark("k3_0")
    r = ops.render_mse(kps, gate, sample_w, student, img_h, img_w, stride, cfg.sigma, grad_scale=grad_scale,
                       want_grad=cfg.want_grad, want_target=cfg.want_target,
                       count_in=count if grad_scale is None else None, loss_weight=cfg.lossWeight)
    if grad_scale is None:
        grad_scale = r["grad_scale"]
    mark("k3_1")
    summary = ops.loss_finalize(r["per_loss"], None, gate.view(B, J))
    out = dict(summary=summary, grad_scale=grad_scale, count=count, grad=r["grad"], target=r["target"],
               gate=gate.view(B, J), kps=kps, enable=sel["enable"].view(B, J), counts=sel["counts"], dist=dist,
               legal=legal, xy=xy, max=mx, idx=idx, per_loss=r["per_loss"], sel=sel)
    out.update(extra)
    return out
