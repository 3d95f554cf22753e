"""Drop-in for the heat-map side of utils/augment.py (AugmentUtils): same classmethod names and
argument meaning.  The forward image augmentation (skimage / cv2, CPU data-loader side) is out of
scope and stays with the reference."""
import math

import numpy as np
import torch

from . import ops


def _to_cuda(t):
    return t if t.is_cuda else t.cuda()


class AugmentUtils:
    @classmethod
    def affine_back2(cls, heatmap, warpmat, isflip, swap_perm=None):
        """utils/augment.py:37-47: back-warp (affine_grid + bilinear grid_sample, zeros padding,
        align_corners=True) and per-sample W mirror, one kernel; returns a new tensor on the
        input's device.  Bit-identical to ATen's CPU kernels.  swap_perm (optional, [J]; default None = the
        reference's behaviour) also exchanges the left/right joint channels of the flipped samples the way
        utils/udaap/transforms.py:20-57 `flip_back` does (ops.swap_perm_from_pairs builds the table)."""
        dev = heatmap.device
        out = ops.warp_materialize(_to_cuda(heatmap.detach()), _to_cuda(warpmat.detach()),
                                   _to_cuda(torch.as_tensor(isflip)), swap_perm=swap_perm)
        return out.to(dev)

    affine_back2_classification = affine_back2      # utils/augment.py:64-74 is the same function

    @classmethod
    def fliplr_back_tensor(cls, flip_output):
        """utils/augment.py:247-252: mirror of the last (W) axis of a 3-D / 4-D tensor (no joint
        swap).  Expressed as affine_back2 with an exact mirror: the identity warp is NOT an exact
        copy in ATen, so the flip is done by the dedicated mirror path with theta unused."""
        if flip_output.ndim not in (3, 4):
            return None                              # the reference falls through and returns None
        x = flip_output if flip_output.ndim == 4 else flip_output.unsqueeze(0)
        dev = x.device
        out = ops.mirror_w(_to_cuda(x.detach().to(torch.float32)))
        out = out.to(dev)
        return out if flip_output.ndim == 4 else out[0]

    @classmethod
    def affine_getWarpmat(cls, angle, scale, matrixRes=[64, 64]):
        """utils/augment.py:159-164: cv2.getRotationMatrix2D(centre, angle, 1/scale) ->
        cv2.invertAffineTransform -> translation zeroed -> float32 [2,3] (host scalars; OpenCV's
        float64 formulas restated, no cv2 dependency)."""
        ang = float(angle) * math.pi / 180.0
        sc = 1.0 / float(scale)
        alpha, beta = math.cos(ang) * sc, math.sin(ang) * sc
        m00, m01, m10, m11 = alpha, beta, -beta, alpha
        D = m00 * m11 - m01 * m10
        D = 1.0 / D if D != 0 else 0.0
        return torch.tensor([[m11 * D, m01 * (-D), 0.0], [m10 * (-D), m00 * D, 0.0]], dtype=torch.float64).float()

    @classmethod
    def view_matrix(cls, center, scale, matrixRes, angle=0):
        """The 3x3 float64 map of utils/udaap/transforms.py:119-148 (get_transform) for one augmented view: crop of
        200*scale around `center` onto matrixRes, rotated by `angle` about the frame centre.  Host arithmetic on
        whatever `scale` / `angle` are (python floats or the 0-d float32 tensors affine_mulKps produces,
        utils/augment.py:28-29), written in the same expression forms so that the dtype rules the reference goes
        through (tensor arithmetic stays float32, python_float / tensor is reciprocal-multiply, numpy sin/cos of a
        float32) apply here too: the matrix is bit-identical, which the integer truncation of the mapped key points
        needs.  Feed the stacked matrices to ops.view_kps."""
        side = 200 * scale
        m = np.zeros((3, 3))
        m[0, 0], m[1, 1], m[2, 2] = float(matrixRes[1]) / side, float(matrixRes[0]) / side, 1
        m[0, 2] = matrixRes[1] * (-float(center[0]) / side + .5)
        m[1, 2] = matrixRes[0] * (-float(center[1]) / side + .5)
        if not angle == 0:
            rad = (-angle) * np.pi / 180
            s, c = np.sin(rad), np.cos(rad)
            spin = np.array([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]])
            to_origin, back = np.eye(3), np.eye(3)
            to_origin[0, 2], to_origin[1, 2] = -matrixRes[1] / 2, -matrixRes[0] / 2
            back[0, 2], back[1, 2] = -to_origin[0, 2], -to_origin[1, 2]
            m = np.dot(back, np.dot(spin, np.dot(to_origin, m)))
        return m

    @classmethod
    def affine_kps(cls, kpsMap, center, scale, matrixRes, angle=0):
        """utils/augment.py:151-156: the visible key points (y > 0) of one sample mapped into the augmented frame;
        returns a new [J,3] tensor on the input's device (the batched form is ops.view_kps)."""
        dev = kpsMap.device
        mat = torch.from_numpy(cls.view_matrix(center, scale, matrixRes, angle)).reshape(1, 1, 3, 3)
        out = ops.view_kps(_to_cuda(kpsMap.detach().to(torch.float32)).unsqueeze(0), mat.cuda(), None, 0.0)
        return out[0, 0].to(dev).to(kpsMap.dtype)
