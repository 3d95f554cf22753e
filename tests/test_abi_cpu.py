"""CPU-only checks of the boundary: the C-ABI library builds/loads and exports every symbol that
include/ubpl_b200.h declares (no compute calls without a GPU), the ctypes table mirrors the header, the
host-side helpers agree with the oracle, and the product path refuses to run on CPU tensors."""
import os
import re

import numpy as np
import pytest
import torch

import ubpl_oracle as O
import ubpl_b200
from ubpl_b200 import _lib, augment, ops, pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "ubpl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ubpl_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = _lib.lib()
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert L.ubpl_version() >= 100


def test_ctypes_table_matches_header():
    hdr = open(os.path.join(ROOT, "include", "ubpl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))
    assert set(_header_functions()) - {"ubpl_last_error"} == set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    x = torch.zeros(1, 2, 8, 8)
    with pytest.raises(_lib.UbplError):
        ops.warp_decode(x, None, None, None)
    with pytest.raises(_lib.UbplError):
        ops.render_targets(torch.zeros(2, 3), 8, 8, 32, 32)
    with pytest.raises(_lib.UbplError):
        ops.ema_flat(torch.zeros(4), torch.zeros(4), 0.5)


@pytest.mark.parametrize("kind", ["f32", "f64", "int"])
def test_decode_coeffs_matches_oracle(kind):
    g = torch.Generator().manual_seed(4)
    B = 16
    center = torch.randint(100, 156, (B, 2), generator=g)
    if kind == "f32":
        scale, sd = (0.8 + torch.rand(B, generator=g)).float(), "f32"
    elif kind == "f64":
        scale, sd = (0.8 + torch.rand(B, generator=g)).double(), "f64"
        center = center.double() + 0.5
    else:
        scale, sd = torch.randint(1, 3, (B,), generator=g), "f32"
    got = ops.decode_coeffs(center, scale, [64, 64]).numpy()
    want = O.decode_coeffs(center.numpy(), scale.numpy(), [64, 64], sd)
    assert np.array_equal(got, want)


def test_host_helpers():
    for ang, sc in [(-20.0, 1 / 1.1), (13.7, 0.8), (0.0, 1.0), (30.0, 1 / 1.6)]:
        assert np.array_equal(augment.AugmentUtils.affine_getWarpmat(ang, sc, [256, 256]).numpy(), O.affine_getWarpmat(ang, sc))
    lab = torch.tensor([True, False, False, True])
    assert torch.equal(pipeline.nega_weights(lab, 0.7), torch.tensor([0.0, 0.7, 0.7, 0.0]))
    import bench
    for c in bench.CONFIGS.values():
        assert bench.algorithmic_bytes_per_sample(c) == 4 * c["H"] * c["W"] * c["J"] * (c["M"] * c["K"] + 2 * c["S"] + 1)
    assert bench.algorithmic_bytes_per_sample(bench.CONFIGS["c2"]) == 2981888


def test_view_matrix_host_helper():
    """augment.AugmentUtils.view_matrix (host side of N1) against the reference's matrices in the golden fixture."""
    from golden_util import load
    g = load("viewkps")
    V, B = g["flips"].shape
    W = int(g["img_w"])
    for v in range(V):
        for b in range(B):
            t = augment.AugmentUtils.view_matrix(g["centers"][v, b].tolist(), torch.tensor(g["scales"][v, b]), [W, W],
                                                 torch.tensor(g["angles"][v, b]))
            np.testing.assert_allclose(t, g["mats"][v, b], rtol=1e-15, atol=1e-15)


def test_workspace_size_helpers():
    """Pure host functions of the ABI: workspace / exchange-buffer sizes (no device needed)."""
    L = _lib.lib()
    # head (128 words) + counts + B*J arrival counters + V*B*J 64-bit hand-off words
    V, B, J = 8, 256, 14
    n = V * B * J
    got = L.ubpl_warp_decode_k2_ws_bytes(V, B, J)
    assert got >= 4 * (128 + J + 2 + B * J + 2 * n) and got % 4 == 0
    assert L.ubpl_warp_decode_k2_ws_bytes(V, 0, J) > 0                      # an empty batch still has the head
    assert L.ubpl_warp_decode_k2_ws_bytes(2 * V, B, J) > got
    # exchange buffer: header + 2 parities x nranks slots x (32-byte slot header + 8 bytes per item)
    assert L.ubpl_p2p_buffer_bytes(8, 4352) == 256 + 2 * 8 * (32 + 8 * 4352)
    assert L.ubpl_p2p_buffer_bytes(17, 10) == 0 and L.ubpl_p2p_buffer_bytes(0, 10) == 0      # 1..16 ranks
    assert L.ubpl_p2p_ranks() == 0 and L.ubpl_p2p_status() == 0            # nothing mapped in this process


def test_grouped_rejects_other_criteria():
    from ubpl_b200 import losses
    with pytest.raises(TypeError):
        losses.grouped(losses.JointDistLoss_mt2(nStack=2), torch.zeros(1, 1, 2, 1, 8, 8), torch.zeros(1, 1, 1, 8, 8))


def test_ema_work_items_cover_every_tensor_once():
    """The chunk / piece tables of the EMA launches (ops.EmaPlan): disjoint work items that cover each tensor exactly."""
    rng = np.random.default_rng(3)
    numels = [1, 3, 1024, 1025, 8192, 8193, 70000] + rng.integers(1, 5000, 20).tolist()
    for size in (ops.EmaPlan.PIECE, ops.EmaPlan.CHUNK, 7):
        which, start = ops.ema_work_items(numels, size)
        assert len(which) == len(start) == sum(-(-n // size) for n in numels)
        seen = [np.zeros(n, np.int32) for n in numels]
        for t, s0 in zip(which, start):
            seen[t][s0:min(s0 + size, numels[t])] += 1
        assert all((v == 1).all() for v in seen)
        assert which == sorted(which)                       # tensor order
    assert ops.EmaPlan.PIECE == 1024 and ops.EmaPlan.CHUNK % ops.EmaPlan.PIECE == 0     # kEmaPiece in csrc/warp_decode.cu
