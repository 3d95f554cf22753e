"""Drop-in for the hot-path functions of utils/process.py (ProcessUtils): decoders, Gaussian
renderer, counters.  Same names, argument meaning and return types; tensors come back on the
device they arrived on (the reference's call sites pass `.cpu()` tensors, projects/MT_UBPL.py:384;
those are moved to the GPU, decoded there and the tiny results returned on the CPU)."""
from itertools import combinations as comb

import torch

from . import ops


def _to_cuda(t):
    return t if t.is_cuda else t.cuda()


class _FeaturesCovFn(torch.autograd.Function):
    """value = mean over (b, n, c) of |cov(f1_row, f2_row)| with both gradients produced by the forward kernel."""

    @staticmethod
    def forward(ctx, inp1, inp2):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        r = ops.features_cov(inp1.detach(), inp2.detach(), want_grad=need)
        ctx.g1, ctx.g2, ctx.shape = r["grad1"], r["grad2"], inp1.shape
        return r["value"]

    @staticmethod
    def backward(ctx, g):
        if ctx.g1 is None:
            return None, None
        scale = g.detach().reshape(1).to(torch.float32).contiguous()
        g1 = ops.scale(ctx.g1, scale).view(ctx.shape) if ctx.needs_input_grad[0] else None
        g2 = ops.scale(ctx.g2, scale).view(ctx.shape) if ctx.needs_input_grad[1] else None
        return g1, g2


class ProcessUtils:
    @classmethod
    def features_cov(cls, inp1, inp2):
        """utils/process.py:19-31: (mean |off-diagonal covariance| of the two views' feature rows, bs*n*c); one
        streaming kernel reads each feature map once and writes both gradients."""
        bs, n, c, h, w = inp1.size()
        return _FeaturesCovFn.apply(inp1, inp2), bs * n * c

    # utils/process.py:53-68 -- scalar helpers on python numbers / 0-d tensors, kept as python
    @classmethod
    def coord_distance(cls, coord1, coord2):
        return ((coord1[0] - coord2[0]) ** 2 + (coord1[1] - coord2[1]) ** 2) ** 0.5

    @classmethod
    def coord_avgDistance(cls, coords):
        dist_sum, dist_n = 0., 0
        for a, b in comb(coords, 2):
            dist_sum += cls.coord_distance(a, b)
            dist_n += 1
        return dist_sum / dist_n                      # ZeroDivisionError for fewer than 2 coords, like the reference

    @classmethod
    def kps_getLabeledCount(cls, kpsGate):
        """utils/process.py:382-383."""
        return int((kpsGate.detach() > 0).sum().item())

    @classmethod
    def kps_fromHeatmap(cls, heatmap, cenMap, scale, res, mode="batch"):
        """utils/process.py:321-327: arg-max decode to image space (+ scores in batch mode)."""
        dev = heatmap.device
        if mode == "single":
            hm = _to_cuda(heatmap.detach().to(torch.float32)).unsqueeze(0)
            dec = ops.decode_coeffs(torch.as_tensor(cenMap).reshape(1, 2), torch.as_tensor(scale).reshape(1), res).cuda()
            return ops.warp_decode(hm, None, None, dec)["xy"][0].to(dev)
        elif mode == "batch":
            hm = _to_cuda(heatmap.detach().to(torch.float32))
            dec = ops.decode_coeffs(cenMap, scale, res).cuda()
            r = ops.warp_decode(hm, None, None, dec, want_idx=False)
            return r["xy"].to(dev), r["max"].cpu()    # the reference's scores are a CPU tensor (numpy round trip)

    @classmethod
    def kps_fromHeatmap_mul(cls, multiOuts, cenMap, scale, res):
        """utils/process.py:330-336 for [K,B,J,H,W] stacks: per-view decode, float32 view means."""
        dev = multiOuts.device
        hm = _to_cuda(multiOuts.detach().to(torch.float32))
        dec = ops.decode_coeffs(cenMap, scale, res).cuda()
        r = ops.warp_decode(hm, None, None, dec, want_idx=False)
        predsMulti, scoresMulti = r["xy"], r["max"]
        predsMean = ops.view_dispersion(predsMulti)["mean"]
        scoresMean = torch.mean(scoresMulti, dim=0)
        return predsMulti.to(dev), predsMean.to(dev), scoresMulti.cpu(), scoresMean.cpu()

    @classmethod
    def kps_fromHeatmap2(cls, heatmap, cenMap, scale, res):
        """utils/process.py:345-379: single-image decoder with the quarter-pixel offset; the
        reference refines joints 0 and 1 only (loop over the coordinate axis, :363) -- kept."""
        dev = heatmap.device
        hm = _to_cuda(heatmap.detach().to(torch.float32)).unsqueeze(0)
        dec = ops.decode_coeffs(torch.as_tensor(cenMap).reshape(1, 2), torch.as_tensor(scale).reshape(1), res).cuda()
        return ops.warp_decode(hm, None, None, dec, refine=1)["xy"][0].to(dev)

    @classmethod
    def kps_heatmap(cls, kpsMap, imgShape, inpRes, outRes, kernelSize=3.0, sigma=1.0):
        """utils/process.py:253-278: Gaussian targets for one image; mutates kpsMap[:, 2] *= vis in
        place and returns (heatmap [J,h,w], kpsMap) like the reference."""
        _, h, w = imgShape
        stride = inpRes / outRes
        sizeH, sizeW = int(h / stride), int(w / stride)
        k = _to_cuda(kpsMap.detach().to(torch.float32))
        hm, kout = ops.render_targets(k[:, :3], sizeH, sizeW, h, w, stride, sigma * kernelSize)
        kpsMap[:, 2] = kout[:, 2].to(device=kpsMap.device, dtype=kpsMap.dtype)
        return hm.to(kpsMap.device).float(), kpsMap

    @classmethod
    def kps_heatmap_mulKps(cls, kpsMapArray, imgShape, inpRes, outRes, kernelSize=3.0, sigma=1.0):
        """utils/process.py:289-318."""
        hms, ks = [], []
        for kpsMap in kpsMapArray:
            hm, k = cls.kps_heatmap(kpsMap, imgShape, inpRes, outRes, kernelSize, sigma)
            hms.append(hm)
            ks.append(k)
        return hms, ks
