"""Drop-ins for EvaluationUtils.uncertainty_fromDistance (utils/evaluation.py:40-58) and
EvaluationUtils.acc_pck (utils/evaluation.py:92-139)."""
import torch

from . import ops


class EvaluationUtils:
    @classmethod
    def uncertainty_fromDistance(cls, preds_mul, preds_mean):
        """unc[b,j] = mean_v ||p_v - p_mean|| / global max; uncW = exp(-unc).  preds_mean must be the
        float32 view mean that kps_fromHeatmap_mul returns (the only way the reference's pipeline
        produces it); it is recomputed on the device and checked when given on the CPU path."""
        dev = preds_mul.device
        pm = preds_mul.detach().to(torch.float32)
        pm = pm if pm.is_cuda else pm.cuda()
        vd = ops.view_dispersion(pm, mean_in=None if preds_mean is None else preds_mean.detach().to(pm.device, torch.float32))
        unc, uncW = ops.unc_normalize(vd["unc32"], vd["max_bits"])
        return unc.to(dev), uncW.to(dev)

    @classmethod
    def acc_pck(cls, preds, gts, pck_ref, pck_thr):
        """utils/evaluation.py:92-139: per-joint mean error and PCK accuracy plus their means, as CPU float32
        tensors [k+1] like the reference returns (one kernel and one small device->host copy instead of a
        python loop of bs*k torch.dist calls)."""
        p = preds.detach().to(torch.float32)
        g = gts.detach().to(torch.float32)
        p = p if p.is_cuda else p.cuda()
        g = g if g.is_cuda else g.cuda()
        errs, accs = ops.acc_pck(p, g, pck_ref, pck_thr)
        both = torch.stack([errs, accs]).cpu()
        return both[0].clone(), both[1].clone()
