"""Drop-in replacements for the reference's heat-map criteria (utils/losses.py), same class
names, constructor arguments, forward signatures and tuple returns (counts are python ints).

    JointMSELoss       utils/losses.py:8-29      JointDistLoss       utils/losses.py:32-53
    JointPseudoLoss3   utils/losses.py:169-210   JointDistLoss_mt2   utils/losses.py:246-286

Each forward is ONE fused CUDA kernel over the student maps (ubpl_dense_mse: loss, score masks and
the gradient in the same pass) plus a tiny reduction, and exactly one device->host sync for the
python-int counts (the reference syncs B*J*S times).  Backward multiplies the stored
gradient by the upstream scalar on the device.  Differences kept deliberately:
  * gradients are produced for the prediction argument (and for preds2/targets when they require
    grad, as minus the prediction gradient); `kpsGate`/`sampleWeight` never get one (the reference
    marks them requires_grad through tools.py:57-62 but nothing consumes those gradients);
  * inputs must be CUDA tensors.
"""
import torch
from torch import nn

from . import ops


def _as5(x, nStack):
    """[B,J,H,W] -> [B,1,J,H,W] when nStack == 1."""
    return x.unsqueeze(1) if nStack == 1 else x


class _DenseLossFn(torch.autograd.Function):
    """loss = sum_{b,s,j} mean_HW (p - t)^2 * coef[b,j] * mask[b,s,j]."""

    @staticmethod
    def forward(ctx, pred5, tgt, coef, mask_mode, thr, gate_for_count, want_scores):
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        r = ops.dense_mse(pred5.detach(), tgt.detach(), coef=coef, mask_mode=mask_mode, thr=thr,
                          want_grad=need_grad, want_scores=want_scores)
        fin = ops.loss_finalize(r["per_loss"], r["mask"] if mask_mode else None, gate_for_count)
        ctx.grad = r["grad"]
        ctx.tgt_shape = tgt.shape
        ctx.pred_shape = pred5.shape
        aux = [t if t is not None else torch.empty(0, device=pred5.device) for t in (r["vmax_p"], r["vmax_t"], r["mask"])]
        ctx.mark_non_differentiable(fin, *aux)        # one call: every call replaces the previous set
        return (fin[0].to(torch.float32), fin) + tuple(aux)

    @staticmethod
    def backward(ctx, g_loss, *unused):
        g = ctx.grad
        if g is None:
            return (None,) * 7
        scale = g_loss.detach().reshape(1).to(torch.float32).contiguous()
        g = ops.scale(g, scale)                     # out of place: the node may be backwarded again (retain_graph)
        g_pred = g if ctx.needs_input_grad[0] else None
        g_tgt = None
        if ctx.needs_input_grad[1]:
            # d/dt = -d/dp, shared targets accumulate over the stacks, M teacher maps share 1/M each
            M = ctx.tgt_shape[0]
            gt = -g / M
            if len(ctx.tgt_shape) == 5:             # [M,B,J,H,W]: shared by the stacks
                gt = gt.sum(1)
            g_tgt = gt.unsqueeze(0).expand(ctx.tgt_shape).contiguous()
        return g_pred, g_tgt, None, None, None, None, None


def _tgt(t):
    """Targets get a gradient only when they are part of a graph (non-leaf); the leaf Variables the
    drivers create with requires_grad=True (projects/tools.py:57-62) are treated as constants."""
    return t if (t.requires_grad and not t.is_leaf) else t.detach()


def _coef(B, J, device, kpsGate, useKPsGate, sampleWeight, useSampleWeight):
    coef = None
    if useKPsGate and kpsGate is not None:
        coef = kpsGate.detach().to(device=device, dtype=torch.float32).reshape(B, J)
    if useSampleWeight and sampleWeight is not None:
        w = sampleWeight.detach().to(device=device, dtype=torch.float32).reshape(B, 1)
        coef = w.expand(B, J) if coef is None else coef * w
    return coef


class JointMSELoss(nn.Module):
    """utils/losses.py:8-29.  forward(preds, gts, kpsGate=None, sampleWeight=None) -> (sum, nStack*count)."""

    def __init__(self, nStack=1, useKPsGate=False, useSampleWeight=False):
        super().__init__()
        self.nStack = nStack
        self.useKPsGate = useKPsGate
        self.useSampleWeight = useSampleWeight

    def forward(self, preds, gts, kpsGate=None, sampleWeight=None):
        p5 = _as5(preds, self.nStack)
        B, S, J = p5.shape[:3]
        coef = _coef(B, J, p5.device, kpsGate, self.useKPsGate, sampleWeight, self.useSampleWeight)
        gate = None if kpsGate is None else kpsGate.detach().to(device=p5.device, dtype=torch.float32).reshape(B, J)
        loss, fin, _, _, _ = _DenseLossFn.apply(p5, _tgt(gts).unsqueeze(0), coef, 0, 0.0, gate, False)
        kpsNum = int(fin[3].item())                       # kps_getLabeledCount (process.py:382-383); B*J if no gate
        return loss, self.nStack * kpsNum


class JointDistLoss(nn.Module):
    """utils/losses.py:32-53.  forward(preds1, preds2, kpsGate=None, sampleWeight=None)."""

    def __init__(self, nStack=1, useKPsGate=False, useSampleWeight=False):
        super().__init__()
        self.nStack = nStack
        self.useKPsGate = useKPsGate
        self.useSampleWeight = useSampleWeight

    def forward(self, preds1, preds2, kpsGate=None, sampleWeight=None):
        p5 = _as5(preds1, self.nStack)
        B, S, J = p5.shape[:3]
        coef = _coef(B, J, p5.device, kpsGate, self.useKPsGate, sampleWeight, self.useSampleWeight)
        gate = None if kpsGate is None else kpsGate.detach().to(device=p5.device, dtype=torch.float32).reshape(B, J)
        t = _as5(_tgt(preds2), self.nStack).unsqueeze(0)  # [1,B,S,J,H,W]: stack s is compared with preds2[:, s]
        loss, fin, _, _, _ = _DenseLossFn.apply(p5, t, coef, 0, 0.0, gate, False)
        return loss, self.nStack * int(fin[3].item())


def _score_mean(v, rows, cnt):
    """mean over the rows with weight > 0 of a [B,S,J] score plane -> [S,J] (losses.py:196-203)."""
    return (v * rows[:, None, None]).sum(0) / cnt


class JointPseudoLoss3(nn.Module):
    """utils/losses.py:169-210.  forward(preds, targets, sampleWeight) ->
    (sum, num_pseudo, num_selected, joint_score_mean[J], scoreThr, scoreThr)."""

    def __init__(self, nStack=1, scoreThr=0.5):
        super().__init__()
        self.nStack = nStack
        self.scoreThr = scoreThr

    def forward(self, preds, targets, sampleWeight):
        p5 = _as5(preds, self.nStack)
        B, S, J = p5.shape[:3]
        targets = _tgt(targets)
        t = targets if self.nStack == 1 else targets[:, :, -1]          # [M,B,J,H,W], losses.py:179
        w = sampleWeight.detach().to(device=p5.device, dtype=torch.float32).reshape(B)
        loss, fin, vp, vt, mask = _DenseLossFn.apply(p5, t, w.reshape(B, 1).expand(B, J), 1, float(self.scoreThr), None, True)
        rows = (w > 0).to(torch.float32)
        cnt = rows.sum()
        host = torch.cat([fin, cnt.reshape(1).double()]).tolist()       # the one device->host sync
        if host[4] == 0:
            raise RuntimeError("stack expects a non-empty TensorList")  # losses.py:201 on an all-labeled batch
        jsm = ((_score_mean(vp, rows, cnt) + _score_mean(vt, rows, cnt)) / 2).mean(0)
        return loss, int(host[1]), int(host[2]), jsm, self.scoreThr, self.scoreThr


class JointDistLoss_mt2(nn.Module):
    """utils/losses.py:246-286.  forward(preds1, preds2, kpsGate=None, sampleWeight=None) ->
    (sum, nStack*count, num_pseudo, num_selected, joint_score_mean[J])."""

    def __init__(self, nStack=1, useKPsGate=False, useSampleWeight=False, scoreThr=0.5):
        super().__init__()
        self.nStack = nStack
        self.useKPsGate = useKPsGate
        self.useSampleWeight = useSampleWeight
        self.scoreThr = scoreThr

    def forward(self, preds1, preds2, kpsGate=None, sampleWeight=None):
        p5 = _as5(preds1, self.nStack)
        B, S, J = p5.shape[:3]
        coef = _coef(B, J, p5.device, kpsGate, self.useKPsGate, sampleWeight, self.useSampleWeight)
        gate = None if kpsGate is None else kpsGate.detach().to(device=p5.device, dtype=torch.float32).reshape(B, J)
        t = _as5(_tgt(preds2), self.nStack).unsqueeze(0)
        loss, fin, _, vt, mask = _DenseLossFn.apply(p5, t, coef, 2, float(self.scoreThr), gate, True)
        w = sampleWeight.detach().to(device=p5.device, dtype=torch.float32).reshape(B)   # AttributeError on None, like :274
        rows = (w > 0).to(torch.float32)
        cnt = rows.sum()
        host = torch.cat([fin, cnt.reshape(1).double()]).tolist()
        if host[4] == 0:
            raise RuntimeError("stack expects a non-empty TensorList")  # losses.py:279
        jsm = _score_mean(vt, rows, cnt).mean(0)
        return loss, self.nStack * int(host[3]), int(host[1]), int(host[2]), jsm


# ------------------------------------------------------------------------------------------------------
# N4: the drivers call a criterion once per (model, augmentation) and sync on every call
# (projects/MT_UBPL.py:246-268).  `grouped` evaluates all G = M*A calls of a JointMSELoss / JointDistLoss in
# ONE loss+gradient kernel and ONE device->host sync; it is what the loops look like once a maintainer
# batches them, not a drop-in: the reference's drivers are unchanged without it.
# ------------------------------------------------------------------------------------------------------
class _GroupedLossFn(torch.autograd.Function):
    """per-group loss sums [G] of a [G*B,S,J,H,W] prediction batch against its targets."""

    @staticmethod
    def forward(ctx, pred5, tgt, coef, gate, G):
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        r = ops.dense_mse(pred5.detach(), tgt.detach(), coef=coef, mask_mode=0, thr=0.0, want_grad=need_grad)
        GB, S, J = r["per_loss"].shape
        B = GB // G
        pl = r["per_loss"].view(G, B, S, J)
        fins = torch.stack([ops.loss_finalize(pl[g], None, None if gate is None else gate.view(G, B, J)[g]) for g in range(G)])
        ctx.grad, ctx.G, ctx.tgt_shape = r["grad"], G, tgt.shape
        ctx.mark_non_differentiable(fins)
        return fins[:, 0].to(torch.float32), fins

    @staticmethod
    def backward(ctx, g_loss, unused):
        g = ctx.grad
        if g is None:
            return None, None, None, None, None
        G = ctx.G
        gv = g.view((G, g.shape[0] // G) + tuple(g.shape[1:]))
        out = torch.empty_like(gv)
        scales = g_loss.detach().to(torch.float32).contiguous()
        for k in range(G):                                   # one tiny scale launch per group, no host sync
            ops.scale(gv[k].contiguous(), scales[k:k + 1], out=out[k])
        out = out.view(g.shape)
        g_pred = out if ctx.needs_input_grad[0] else None
        g_tgt = None
        if ctx.needs_input_grad[1]:
            gt = -out                                        # d/dt = -d/dp
            if len(ctx.tgt_shape) == 5:                      # [1,G*B,J,H,W]: one target shared by the stacks
                gt = gt.sum(1)
            g_tgt = gt.reshape(ctx.tgt_shape)
        return g_pred, g_tgt, None, None, None


class _GroupedPseudoFn(torch.autograd.Function):
    """JointPseudoLoss3 over G groups at once: per-group masked loss sums [G] of a [G*B,S,J,H,W] prediction batch
    against the mean of the Mt teacher maps [Mt,G*B,J,H,W] (last stack), score masks and gradient in the same pass."""

    @staticmethod
    def forward(ctx, pred5, tgt, coef, thr, G):
        need_grad = ctx.needs_input_grad[0]
        r = ops.dense_mse(pred5.detach(), tgt.detach(), coef=coef, mask_mode=1, thr=thr, want_grad=need_grad, want_scores=True)
        GB, S, J = r["per_loss"].shape
        B = GB // G
        pl, mk = r["per_loss"].view(G, B, S, J), r["mask"].view(G, B, S, J)
        fins = torch.stack([ops.loss_finalize(pl[g], mk[g], None) for g in range(G)])
        ctx.grad, ctx.G = r["grad"], G
        ctx.mark_non_differentiable(fins, r["vmax_p"], r["vmax_t"])
        return fins[:, 0].to(torch.float32), fins, r["vmax_p"], r["vmax_t"]

    @staticmethod
    def backward(ctx, g_loss, *unused):
        g = ctx.grad
        if g is None:
            return None, None, None, None, None
        G = ctx.G
        gv = g.view((G, g.shape[0] // G) + tuple(g.shape[1:]))
        out = torch.empty_like(gv)
        scales = g_loss.detach().to(torch.float32).contiguous()
        for k in range(G):                                   # one tiny scale launch per group, no host sync
            ops.scale(gv[k], scales[k:k + 1], out=out[k])
        return out.view(g.shape), None, None, None, None


def _grouped_pseudo(criterion, preds, targets, sampleWeight):
    """All G calls `JointPseudoLoss3(preds[g], targets[:, g], sampleWeight)` (utils/losses.py:169-210; the epc loop of
    projects/MT_UBPL.py:270-298 makes M * K of them per step, each with O(B*J*S) host syncs in the reference) in ONE
    loss+gradient kernel and ONE device->host sync.  preds [G,B,(S,)J,H,W]; targets [Mt,G,B,(S,)J,H,W] -- the teacher
    stacks each group is compared with, treated as constants like the drivers' `.clone().detach()`; sampleWeight [B,1]
    (the same for every group).  Returns (sums [G] with autograd, num_pseudo list, num_selected list,
    joint_score_mean [G,J], scoreThr, scoreThr)."""
    nS = criterion.nStack
    G, B = preds.shape[0], preds.shape[1]
    p5 = _as5(preds.reshape((G * B,) + tuple(preds.shape[2:])), nS)
    S, J = p5.shape[1], p5.shape[2]
    t = targets.detach()
    t = t if nS == 1 else t[:, :, :, -1]                          # [Mt,G,B,J,H,W]: the last stack, losses.py:179
    Mt = t.shape[0]
    if t.stride(1) == B * t.stride(2):
        t = t.as_strided((Mt, G * B) + tuple(t.shape[3:]), (t.stride(0), t.stride(2)) + tuple(t.stride()[3:]))
    else:
        t = t.reshape((Mt, G * B) + tuple(t.shape[3:]))
    w = sampleWeight.detach().to(device=p5.device, dtype=torch.float32).reshape(B)
    coef = w.reshape(1, B, 1).expand(G, B, J).reshape(G * B, J).contiguous()
    sums, fins, vp, vt = _GroupedPseudoFn.apply(p5, t, coef, float(criterion.scoreThr), G)
    rows = (w > 0).to(torch.float32)
    cnt = rows.sum()
    host = torch.cat([fins.reshape(-1), cnt.reshape(1).double()]).tolist()      # the one device->host sync
    if host[-1] == 0:
        raise RuntimeError("stack expects a non-empty TensorList")              # losses.py:201 on an all-labeled batch
    vp, vt = vp.view(G, B, S, J), vt.view(G, B, S, J)
    jsm = torch.stack([((_score_mean(vp[g], rows, cnt) + _score_mean(vt[g], rows, cnt)) / 2).mean(0) for g in range(G)])
    n_pseudo = [int(host[4 * g + 1]) for g in range(G)]
    n_sel = [int(host[4 * g + 2]) for g in range(G)]
    return sums, n_pseudo, n_sel, jsm, criterion.scoreThr, criterion.scoreThr


def grouped(criterion, preds, targets, kpsGate=None, sampleWeight=None):
    """All G calls `criterion(preds[g], targets[g], kpsGate[g], sampleWeight)` of a JointMSELoss / JointDistLoss
    (utils/losses.py:8-53) at once; a JointPseudoLoss3 goes to _grouped_pseudo (its targets are [Mt,G,B,...]).  preds [G,B,(S,)J,H,W]; targets [G,B,J,H,W] (JointMSELoss: one target per
    sample, shared by the stacks) or [G,B,(S,)J,H,W] (JointDistLoss); kpsGate [G,B,J] or None; sampleWeight [B,1]
    (the same for every group, as in the drivers).  Returns (sums [G] float32 with autograd, counts: list of G
    python ints) -- the tuples the G separate calls would have returned, with one kernel and one sync."""
    if isinstance(criterion, JointPseudoLoss3):
        return _grouped_pseudo(criterion, preds, targets, sampleWeight)
    if not isinstance(criterion, (JointMSELoss, JointDistLoss)):
        raise TypeError("grouped() batches JointMSELoss / JointDistLoss / JointPseudoLoss3 calls")
    nS = criterion.nStack
    G, B = preds.shape[0], preds.shape[1]
    p5 = preds.reshape((G * B,) + tuple(preds.shape[2:]))
    p5 = _as5(p5, nS)
    S, J = p5.shape[1], p5.shape[2]
    t = _tgt(targets)
    t = t.reshape((G * B,) + tuple(t.shape[2:]))
    if isinstance(criterion, JointMSELoss):
        t = t.unsqueeze(0)                                   # [1,G*B,J,H,W]: shared by the stacks
    else:
        t = _as5(t, nS).unsqueeze(0)                         # [1,G*B,S,J,H,W]
    gate = None if kpsGate is None else kpsGate.detach().to(device=p5.device, dtype=torch.float32).reshape(G * B, J)
    coef = None
    if criterion.useKPsGate and gate is not None:
        coef = gate
    if criterion.useSampleWeight and sampleWeight is not None:
        w = sampleWeight.detach().to(device=p5.device, dtype=torch.float32).reshape(1, B, 1).expand(G, B, J).reshape(G * B, J)
        coef = w if coef is None else coef * w
    coef = None if coef is None else coef.contiguous()
    sums, fins = _GroupedLossFn.apply(p5, t, coef, gate, G)
    counts = [nS * int(c) for c in fins[:, 3].tolist()]      # the one device->host sync
    return sums, counts
