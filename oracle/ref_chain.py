"""The pseudo-label chain of SURVEY.md section 3.4 / 8d run through the REFERENCE'S OWN functions (imported
unmodified by oracle/ref_import.py)  --  TEST INFRASTRUCTURE and CPU BASELINE, NOT PRODUCT CODE.

Used by tests/test_oracle_vs_reference.py (pins oracle/ubpl_oracle.py::pseudo_label_chain against it) and by
bench.py's `cpu_baseline` / `--impl reference` legs (kind "reference"): the reference's implementation of the path
timed on the GPU box's host cores.  The product package never imports it.

Chain (reference file:line):
  affine_back2 per view                         utils/augment.py:37-47
  kps_fromHeatmap_mul per teacher               utils/process.py:330-336 (final_preds, transform_preds, transform)
  M = 1: uncertainty_fromDistance               utils/evaluation.py:40-58
         (fixed rule unc <= 1-exp(-3*distThrMax/5) on the mean view distance, business.py:375-376, 253)
  M = 2 or quantile: assess_pseudo_unc(2) + filter_pseudo2      utils/business.py:16-35,109-217
  kps_heatmap per sample                        utils/process.py:253-278
  JointMSELoss(nStack, useKPsGate, useSampleWeight) forward + backward, weight*sum/n      utils/losses.py:8-29,
                                                projects/MT_UBPL.py:258-268; weights projects/tools.py:24-31
  update_ema_variables                          utils/parameters.py:4-8
"""
import math
import types

import numpy as np
import torch


def _args(K, J, distThrMax, reliableThr, reliablePCT, reliableDistMin, pseudoWeight):
    return types.SimpleNamespace(pck_ref=[0, 1], pck_thr=0.2, br_inferAugNum=K, reliableThr=reliableThr,
                                 reliablePCT=reliablePCT, reliableDistMin=reliableDistMin, kpsCount=J,
                                 distThrMax=distThrMax, pseudoWeight=pseudoWeight, device="cpu")


def reference_chain(ref, d, select="fixed", distThrMax=1.0, reliableThr=0.0, reliablePCT=0.5, reliableDistMin=1.0,
                    pseudoWeight=1.0, lossWeight=1.0, stride=4.0, sigma=3.0, tools=None, want_grad=True):
    """d: a synth.make_batch dict of CPU tensors.  Returns numpy outputs named like
    ubpl_oracle.pseudo_label_chain's (idx, max, xy, kps, enable, gate, target, loss, count, grad)."""
    teacher, student = d["teacher"], d["student"]
    M, K, B, J, H, W = teacher.shape
    S = student.shape[1]
    res = [H, W]
    args = _args(K, J, distThrMax, reliableThr, reliablePCT, reliableDistMin, pseudoWeight)
    back = torch.stack([torch.stack([ref.aug.affine_back2(teacher[m, v], d["theta"][v], d["flip"][v])
                                     for v in range(K)]) for m in range(M)])
    dec = [ref.proc.kps_fromHeatmap_mul(back[m], d["center"], d["scale"], res) for m in range(M)]
    maxv, idx = torch.max(back.reshape(M, K, B, J, -1), -1)
    xy = torch.stack([x[0] for x in dec])
    ids = ["im%d" % b for b in range(B)]
    gt = torch.cat([d["base_xy"] * stride + 1, torch.ones(B, J, 1)], -1)
    extra = {}
    if M == 1:
        kps = dec[0][1]
        unc_n, _ = ref.eval.uncertainty_fromDistance(dec[0][0], dec[0][1])           # normalised by its global maximum
        legal = (xy[0] >= 0).all(0).all(-1)
        if select == "fixed":
            # the fixed rule needs the un-normalised mean view distance: the same python-float arithmetic
            # (utils/evaluation.py:61-62 `_calDist_fromCoords`) over the same lists
            pm, pb = dec[0][0].double().tolist(), dec[0][1].double().tolist()
            dist = torch.tensor([[sum(((pm[v][b][j][0] - pb[b][j][0]) ** 2 + (pm[v][b][j][1] - pb[b][j][1]) ** 2) ** 0.5
                                      for v in range(K)) / K for j in range(J)] for b in range(B)], dtype=torch.float64)
            dist = torch.where(legal, dist, torch.full_like(dist, 999.0))
            thr = 1 - math.exp(-(distThrMax * 3) / 5)
            unc = torch.tensor([1 - math.exp(-x / 5) for x in dist.reshape(-1).tolist()], dtype=torch.float64).reshape(B, J)
            enable = legal & (unc <= thr)
        else:
            # global quantile on the single teacher's dispersion: the records of assess_pseudo_unc with the
            # dispersion in the extDist field, through filter_pseudo2
            recs = ref.bus.assess_pseudo_unc(ids, gt, [kps], args)[0]
            pm, pb = dec[0][0].double().tolist(), dec[0][1].double().tolist()
            for r_ in recs:
                b, j = ids.index(r_["imageID"]), r_["kIdx"]
                dd = sum(((pm[v][b][j][0] - pb[b][j][0]) ** 2 + (pm[v][b][j][1] - pb[b][j][1]) ** 2) ** 0.5 for v in range(K)) / K
                r_["extDist"] = dd if bool(legal[b, j]) else 999
                r_["coord_legal"] = 1.0 if bool(legal[b, j]) else 0.0
            sel, counts, errs, accs, thr = ref.bus.filter_pseudo2(recs, args)
            enable = torch.zeros(B, J, dtype=torch.bool)
            for it in sel:
                enable[ids.index(it["imageID"]), int(it["kpID"].split("_")[-1])] = bool(it["enable"])
            dist = None
            extra["thr"] = thr
    else:
        pm1, pm2 = dec[0][0], dec[1][0]
        p1, p2 = dec[0][1], dec[1][1]
        pmean = ref.bus.preds_mean(p1, p2)
        pseudo, _, _ = ref.bus.assess_pseudo_unc2(ids, gt, [p1, p2, pmean], [list(pm1), list(pm2)], args)
        kps = torch.tensor([[pseudo[b * J + j]["coord"] for j in range(J)] for b in range(B)], dtype=torch.float64).float()
        dist = torch.tensor([[float(pseudo[b * J + j]["extDist"]) for j in range(J)] for b in range(B)], dtype=torch.float64)
        legal = torch.tensor([[pseudo[b * J + j]["coord_legal"] > 0 for j in range(J)] for b in range(B)])
        if select == "fixed":
            thr = 1 - math.exp(-(distThrMax * 3) / 5)
            unc = torch.tensor([1 - math.exp(-x / 5) for x in dist.reshape(-1).tolist()], dtype=torch.float64).reshape(B, J)
            enable = legal & (unc <= thr)
        else:
            sel, counts, errs, accs, thr = ref.bus.filter_pseudo2(pseudo, args)
            enable = torch.zeros(B, J, dtype=torch.bool)
            for it in sel:
                enable[ids.index(it["imageID"]), int(it["kpID"].split("_")[-1])] = bool(it["enable"])
            extra["thr"] = thr
    img_h, img_w = int(H * stride), int(W * stride)
    targets, gates = [], []
    for b in range(B):
        k3 = torch.cat([kps[b].float(), enable[b].float()[:, None]], -1)
        hm, kout = ref.proc.kps_heatmap(k3, (3, img_h, img_w), img_h, H, kernelSize=sigma, sigma=1.0)
        targets.append(hm)
        gates.append(kout[:, 2])
    target, gate = torch.stack(targets), torch.stack(gates)
    w = torch.where(d["islabeled"].bool(), torch.zeros(B), torch.full((B,), float(pseudoWeight))).unsqueeze(-1)
    crit = ref.losses.JointMSELoss(nStack=S, useKPsGate=True, useSampleWeight=True)
    preds = student.clone().requires_grad_(want_grad)
    loss_sum, n = crit(preds, target, gate, w)
    loss = lossWeight * loss_sum / n if n > 0 else lossWeight * loss_sum
    grad = None
    if want_grad:
        loss.backward()
        grad = preds.grad.numpy()
    out = dict(idx=idx.numpy(), max=maxv.numpy(), xy=xy.numpy(), kps=kps.numpy(), enable=enable.numpy(),
               gate=gate.numpy(), target=target.numpy(), loss=float(loss.detach()), count=int(n), grad=grad,
               dist=None if dist is None else dist.numpy(), legal=legal.numpy())
    out.update(extra)
    return out


def make_hourglass_pair(ref_root_ns, J, nStack=2, seed=0):
    """Two reference StackedHourglass models (student, EMA teacher) on the CPU -- the class itself, not the
    `.cuda()` factory of models/pose/pose_model.py:8."""
    import importlib
    hg = importlib.import_module("models.pose.hourglass")
    torch.manual_seed(seed)
    model = hg.StackedHourglass(J, nStack, "AvgPool")
    ema = hg.StackedHourglass(J, nStack, "AvgPool")
    for p in ema.parameters():
        p.detach_()
    return model, ema


def reference_ema(ref, model, ema, epo=3, ema_decay=0.999):
    """utils/parameters.py:4-8 on the reference's own models."""
    ref.parameters.update_ema_variables(model, ema, types.SimpleNamespace(epo=epo, ema_decay=ema_decay))
