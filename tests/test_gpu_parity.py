"""GPU parity tests: the CUDA path, called through the C ABI (ubpl_b200.ops -> ctypes ->
libubpl_b200.so), against the oracle on seeded inputs and against the committed golden vectors
(reference outputs).  Bars: indices, masks, integer coordinates bit-exact; the warped maps
bit-exact (same float op order as ATen's CPU kernels); losses/targets/gradients within 1e-5
relative (tolerance stated at each assert)."""
import numpy as np
import pytest
import torch

import ubpl_oracle as O
from golden_util import load

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ubpl_b200
    from ubpl_b200 import ops as _ops
    return _ops


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def npy(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# K1: warp
# ------------------------------------------------------------------------------------------------
def test_warp_materialize_golden_bit_exact(ops):
    g = load("chain_mt")
    t = g["teacher"]
    for v in range(t.shape[1]):
        got = ops.warp_materialize(cu(t[0, v]), cu(g["theta"][v]), cu(g["flip"][v]))
        assert np.array_equal(npy(got), g["back"][0, v])


@pytest.mark.parametrize("shape", [(64, 64), (32, 48), (128, 128), (17, 23), (8, 8)])
def test_warp_materialize_vs_oracle(ops, shape):
    H, W = shape
    rng = np.random.default_rng(H * 131 + W)
    x = rng.standard_normal((5, 3, H, W)).astype(np.float32)
    th = (rng.standard_normal((5, 2, 3)) * 0.6).astype(np.float32)
    th[0] = [[1, 0, 0], [0, 1, 0]]                 # identity (not an exact copy in ATen either)
    th[1] = [[0, 0, 0], [0, 0, 0]]                 # singular
    fl = np.array([1, 0, 1, 0, 1], np.uint8)
    want = O.affine_back2(x, th, fl)
    got = ops.warp_materialize(cu(x), cu(th), cu(fl))
    assert np.array_equal(npy(got), want)


def _decode_oracle(maps, theta, flip, center, scale, sd="f32"):
    """maps [V,B,J,H,W] -> idx, max, xy via the oracle (back-warp per view, argmax, transform)."""
    V, B, J, H, W = maps.shape
    back = np.stack([O.affine_back2(maps[v], theta[v], flip[v]) for v in range(V)])
    val, idx = O.argmax_first(back)
    xy = np.stack([O.final_preds(back[v], center, scale, [H, W], sd) for v in range(V)])
    return idx, val, xy


@pytest.mark.parametrize("name", ["chain_mt", "chain_dual"])
def test_warp_decode_golden(ops, name):
    g = load(name)
    t = g["teacher"]
    M, K, B, J, H, W = t.shape
    dec = ops.decode_coeffs(torch.as_tensor(g["center"]), torch.as_tensor(g["scale"]), [H, W]).cuda()
    for m in range(M):
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        r = ops.warp_decode(cu(t[m]), cu(g["theta"]), cu(g["flip"]), dec, stats=stats)
        assert np.array_equal(npy(r["idx"]).astype(np.int64), g["argmax_idx"][m])       # bit-exact indices
        assert np.array_equal(npy(r["max"]), g["max_val"][m])                           # bit-exact scores
        assert np.array_equal(npy(r["xy"]), g["preds_multi"][m])                        # integer coords
        assert int(stats[2]) == K * B * J


def test_warp_decode_vs_oracle_batch(ops):
    import ubpl_b200
    from ubpl_b200 import synth
    d = synth.make_batch(B=12, K=4, J=14, M=1, S=1, seed=4242)
    t = d["teacher"][0].numpy()
    idx, val, xy = _decode_oracle(t, d["theta"].numpy(), d["flip"].numpy(), d["center"].numpy(), d["scale"].numpy())
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64]).cuda()
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = ops.warp_decode(d["teacher"][0].cuda(), d["theta"].cuda(), d["flip"].cuda(), dec, stats=stats)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)
    assert np.array_equal(npy(r["xy"]), xy)
    n_maps, n_slow = int(stats[2]), int(stats[0])
    assert n_maps == 4 * 12 * 14
    # the pruned path must carry the bulk of the maps (all-negative maps take the exhaustive path)
    assert n_slow < 0.25 * n_maps, (n_slow, n_maps)


def test_warp_decode_edge_cases(ops):
    H = W = 64
    rng = np.random.default_rng(7)
    maps = np.zeros((3, 2, 8, H, W), np.float32)
    maps[:, :, 0] = 1.0                                    # constant: every tie -> first index
    maps[:, :, 1] = -1.0                                   # all negative: max is 0 from the zero padding or negative
    maps[:, :, 2, 10, 20] = 0.5
    maps[:, :, 2, 40, 3] = 0.5                             # exact tie between two far texels
    maps[:, :, 3] = rng.standard_normal((H, W)) * 1e-3     # noise only, tiny positive max
    maps[:, :, 4, 0, 0] = 2.0                              # peak in the corner
    maps[:, :, 5, 63, 63] = 2.0
    maps[:, :, 6] = rng.standard_normal((H, W))
    maps[:, :, 6, 30, 30] = np.nan                         # NaN propagates like torch.max
    maps[:, :, 7] = np.abs(rng.standard_normal((H, W))) + 5.0
    th = np.zeros((3, 2, 2, 3), np.float32)
    th[0, 0] = [[1, 0, 0], [0, 1, 0]]
    th[0, 1] = [[0.7, 0.3, 0.05], [-0.3, 0.7, -0.02]]
    th[1, 0] = [[1.5, 0.4, 0], [-0.4, 1.5, 0]]              # zoom out: zero padding visible
    th[1, 1] = [[0, 0, 0], [0, 0, 0]]                       # singular
    th[2, 0] = [[0.3, 0, 0.5], [0, 0.3, -0.5]]              # strong zoom in
    th[2, 1] = [[-0.9, 0.1, 0], [0.1, 0.9, 0]]              # reflection
    fl = np.array([[0, 1], [1, 0], [1, 1]], np.uint8)
    center = np.full((2, 2), 128, np.int64)
    scale = np.array([1.28, 1.0], np.float32)
    idx, val, xy = _decode_oracle(maps, th, fl, center, scale)
    dec = ops.decode_coeffs(torch.as_tensor(center), torch.as_tensor(scale), [H, W]).cuda()
    r = ops.warp_decode(cu(maps), cu(th), cu(fl), dec)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val, equal_nan=True)
    assert np.array_equal(npy(r["xy"]), xy)


def test_warp_decode_non_positive_maps(ops):
    """All-negative / non-positive maps under many transforms: the maximum is 0 at the first pixel that
    samples only padding, or a negative interior value when the frame stays inside the source."""
    rng = np.random.default_rng(23)
    V, B, J, H, W = 6, 8, 6, 64, 64
    maps = -(np.abs(rng.standard_normal((V, B, J, H, W))) * 0.02 + 1e-3).astype(np.float32)
    maps[:, :, 1] *= 50.0
    maps[:, :, 2, 5:9, 7:11] = 0.0                         # maximum exactly 0 inside the map
    maps[:, :, 3] = -1.0                                   # constant negative
    maps[:, :, 4, 20, 20] = -1e-30                         # tiny magnitudes
    maps[:, :, 5] = np.minimum(maps[:, :, 5], -0.5) + 0.4999
    ang = rng.uniform(-0.8, 0.8, (V, B))
    sc = rng.uniform(0.5, 1.7, (V, B))
    th = np.zeros((V, B, 2, 3), np.float32)
    th[..., 0, 0] = np.cos(ang) * sc; th[..., 0, 1] = np.sin(ang) * sc
    th[..., 1, 0] = -np.sin(ang) * sc; th[..., 1, 1] = np.cos(ang) * sc
    th[..., 2] = rng.uniform(-0.3, 0.3, (V, B, 2))
    th[0, 0] = [[1, 0, 0], [0, 1, 0]]
    th[0, 1] = [[1, 0, 2.0 / 63], [0, 1, 0]]                # shifted by exactly one texel: ix = W at the last column
    th[0, 2] = [[1, 0, -2.0 / 63], [0, 1, -2.0 / 63]]       # ix = -1 exactly at column 0
    th[0, 3] = [[0.5, 0, 0], [0, 0.5, 0]]                   # fully inside
    fl = rng.integers(0, 2, (V, B)).astype(np.uint8)
    idx, val, xy = _decode_oracle(maps, th, fl, np.full((B, 2), 128), np.full(B, 1.28, np.float32))
    dec = ops.decode_coeffs(torch.full((B, 2), 128), torch.full((B,), 1.28), [H, W]).cuda()
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = ops.warp_decode(cu(maps), cu(th), cu(fl), dec, stats=stats)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)
    assert np.array_equal(npy(r["xy"]), xy)


def test_warp_decode_strided_and_odd_shapes(ops):
    rng = np.random.default_rng(11)
    # the reference slices outs_ema[m, a, :, -1]: the S axis is skipped
    full = rng.standard_normal((2, 5, 3, 6, 32, 48)).astype(np.float32)     # [V,B,S,J,H,W]
    th = (rng.standard_normal((2, 5, 2, 3)) * 0.5).astype(np.float32)
    fl = rng.integers(0, 2, (2, 5)).astype(np.uint8)
    center = np.full((5, 2), 96, np.int64)
    scale = np.full((5,), 1.0, np.float32)
    sl = full[:, :, -1]
    idx, val, xy = _decode_oracle(np.ascontiguousarray(sl), th, fl, center, scale)
    dec = ops.decode_coeffs(torch.as_tensor(center), torch.as_tensor(scale), [32, 48]).cuda()
    r = ops.warp_decode(cu(full)[:, :, -1], cu(th), cu(fl), dec)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)
    assert np.array_equal(npy(r["xy"]), xy)
    # 17x23 planes: 391 texels, not a multiple of 4 -> the non-bulk staging path
    odd = rng.standard_normal((1, 3, 4, 17, 23)).astype(np.float32)
    th2 = (rng.standard_normal((1, 3, 2, 3)) * 0.5).astype(np.float32)
    fl2 = np.array([[1, 0, 1]], np.uint8)
    idx, val, _ = _decode_oracle(odd, th2, fl2, np.full((3, 2), 40), np.ones(3, np.float32))
    r = ops.warp_decode(cu(odd), cu(th2), cu(fl2), None)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)


def test_plain_decode_golden(ops):
    g = load("decode")
    hm = cu(g["hm"])
    for k in ("f32_128", "f32_one", "int_one", "f32_rand", "f64_rand"):
        dec = ops.decode_coeffs(torch.as_tensor(g[k + "_center"]), torch.as_tensor(g[k + "_scale"]), [64, 64]).cuda()
        r = ops.warp_decode(hm, None, None, dec)
        assert np.array_equal(npy(r["xy"]), g[k + "_preds"]), k
        assert np.array_equal(npy(r["max"]), g[k + "_scores"], equal_nan=True), k
    # quarter-offset decoder, bug parity with utils/process.py:363 (joints 0 and 1 only)
    dec = ops.decode_coeffs(torch.tensor([[128, 128]]), torch.tensor([1.28]), [64, 64]).cuda()
    r = ops.warp_decode(cu(g["q_hm"])[None], None, None, dec, refine=1)
    assert np.array_equal(npy(r["xy"])[0], g["q_preds"])
    want_all = O.kps_fromHeatmap2(g["q_hm"], np.array([128, 128]), np.array(1.28, np.float32), [64, 64], refine="all")
    r = ops.warp_decode(cu(g["q_hm"])[None], None, None, dec, refine=2)
    assert np.array_equal(npy(r["xy"])[0], want_all)


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def test_view_dispersion_golden(ops):
    g = load("chain_mt")
    pm = cu(g["preds_multi"][0])
    vd = ops.view_dispersion(pm)
    assert np.array_equal(npy(vd["mean"]), g["preds_mean"][0])
    unc, uncW = ops.unc_normalize(vd["unc32"], vd["max_bits"])
    np.testing.assert_allclose(npy(unc), g["unc"], rtol=1e-6)        # libm pow vs IEEE sqrt: <= 1 ulp of float64
    np.testing.assert_allclose(npy(uncW), g["uncW"], rtol=1e-6)


def test_dual_assess_and_quantile_golden(ops):
    g = load("chain_dual")
    B, J = g["dual_extDist"].shape
    ad = ops.assess_dual(cu(g["dual_p1"]), cu(g["dual_p2"]), cu(g["dual_pmean"]), cu(g["preds_multi"][0]), cu(g["preds_multi"][1]))
    for k_ref, k in (("coord_legal", "legal"), ("intDist1", "intDist1"), ("intDist2", "intDist2"), ("extDist", "extDist"),
                     ("coord_w1", "w1"), ("coord_w2", "w2"), ("coord", "coord")):
        assert np.array_equal(npy(ad[k]), g["dual_" + k_ref]), k          # float64, bit-exact
    assert int(ad["zero_div"]) == 0
    for pct in (25, 50, 90):
        s = ops.select_quantile(ad["extDist"], ad["legal"], J, 0.0, pct / 100.0, 1.0)
        assert float(s["thr"]) == float(g["filt%02d_thr" % pct])
        assert np.array_equal(npy(s["enable"]).reshape(B, J).astype(np.int32), g["filt%02d_enable" % pct])   # masks bit-exact
        assert np.array_equal(npy(s["reliability"]).reshape(B, J), g["filt%02d_reliability" % pct])
        assert np.array_equal(npy(s["counts"]).astype(np.int64), g["filt%02d_counts" % pct])


def test_quantile_select_random(ops):
    rng = np.random.default_rng(3)
    for n, J in ((1, 1), (7, 7), (1000, 10), (34816, 17)):
        dist = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4          # many exact ties
        dist[rng.random(n) < 0.2] = 999.0
        legal = (rng.random(n) < 0.9).astype(np.float64)
        for pct, rthr in ((0.5, 0.0), (0.1, 0.0), (0.99, 0.0), (0.5, 0.8), (0.0, 0.0), (1.0, 0.0)):
            rel, thr, en = O.filter_dual(dist, legal, rthr, pct, 1.0)
            for backend in (None, ops._CudaSelectBackend()):      # one-launch path, then the multi-kernel (multi-GPU) path
                s = ops.select_quantile(cu(dist), cu(legal), J, rthr, pct, 1.0, backend=backend)
                assert float(s["thr"]) == thr, (n, pct)
                assert np.array_equal(npy(s["enable"]).astype(bool), en)
                assert np.array_equal(npy(s["reliability"]), rel)
                assert int(s["counts"][-1]) == int(en.sum())


def test_quantile_select_nccl_single_rank(ops):
    """The NCCL-fused selector (the multi-GPU path) on a 1-rank communicator: same threshold and masks."""
    import ctypes
    from ubpl_b200 import _lib
    buf = (ctypes.c_char * 128)()
    _lib.call("ubpl_nccl_unique_id", ctypes.cast(buf, ctypes.c_void_p))
    _lib.call("ubpl_nccl_init", ctypes.cast(buf, ctypes.c_void_p), 1, 0)
    try:
        rng = np.random.default_rng(9)
        for n, J in ((7, 7), (4352, 17)):
            dist = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4
            dist[rng.random(n) < 0.2] = 999.0
            legal = (rng.random(n) < 0.9).astype(np.float64)
            for pct in (0.5, 0.1, 0.99):
                rel, thr, en = O.filter_dual(dist, legal, 0.0, pct, 1.0)
                s = ops.select_quantile_nccl(cu(dist), cu(legal), J, int((n - 1) * pct), 0.0, 1.0)
                assert float(s["thr"]) == thr
                assert np.array_equal(npy(s["enable"]).astype(bool), en)
                assert np.array_equal(npy(s["reliability"]), rel)
    finally:
        _lib.call("ubpl_nccl_destroy")


def test_select_fixed(ops):
    rng = np.random.default_rng(5)
    dist = np.concatenate([rng.gamma(2.0, 2.0, 500), [3.0, 9.0, 0.0, 999.0]])
    legal = (rng.random(dist.size) < 0.9).astype(np.float64)
    for thr in (1.0, 3.0):
        want = (legal > 0) & (np.array([O.unc_value(d) for d in dist]) <= O.unc_value(thr * 3))
        s = ops.select_fixed(cu(dist), cu(legal), 4, thr)
        assert np.array_equal(npy(s["enable"]).astype(bool), want)
        assert int(s["counts"][-1]) == int(want.sum())


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
def test_render_targets_golden(ops):
    g = load("render")
    hm, kout = ops.render_targets(cu(g["kps"]), 64, 64, 256, 256)
    hm = npy(hm)
    assert np.array_equal(hm == 0, g["heatmap"] == 0)                      # identical support (the 0.01 cut)
    np.testing.assert_allclose(hm, g["heatmap"], rtol=RTOL, atol=0)        # 1e-5 relative
    assert np.array_equal(npy(kout), g["kps_out"])


def test_render_mse_golden(ops):
    g = load("chain_mt")
    B, S, J, H, W = g["student"].shape
    kps = cu(g["preds_mean"][0])
    r = ops.render_mse(kps, None, cu(g["weight"]), cu(g["student"]), 256, 256)
    assert np.array_equal(npy(r["gate_out"]), g["gate"])
    tgt = npy(r["target"])
    assert np.array_equal(tgt == 0, g["target"] == 0)
    np.testing.assert_allclose(tgt, g["target"], rtol=RTOL)
    fin = npy(ops.loss_finalize(r["per_loss"], None, r["gate_out"]))
    np.testing.assert_allclose(fin[0], float(g["mse_loss"]), rtol=RTOL)
    assert int(fin[3]) * S == int(g["mse_count"])
    np.testing.assert_allclose(npy(r["grad"]), g["mse_grad"], rtol=RTOL, atol=1e-9)


def test_dense_losses_golden(ops):
    g = load("losses")
    p = cu(g["student"])
    B, S, J, H, W = p.shape
    t = cu(g["targets"])[:, :, -1]                      # targets[:, :, -1]: strided view, like losses.py:179
    r = ops.dense_mse(p, t, coef=cu(g["nega"]).expand(B, J), mask_mode=1, thr=0.8, want_scores=True)
    fin = npy(ops.loss_finalize(r["per_loss"], r["mask"], None))
    np.testing.assert_allclose(fin[0], float(g["p3_loss"]), rtol=RTOL)
    assert (int(fin[1]), int(fin[2])) == (int(g["p3_num_pseudo"]), int(g["p3_num_selected"]))
    np.testing.assert_allclose(npy(r["grad"]), g["p3_grad"], rtol=RTOL, atol=1e-9)
    o = O.joint_pseudo3(g["student"], g["targets"], g["nega"], 2, 0.8)
    assert np.array_equal(npy(r["mask"]), o["mask"])                                   # masks bit-exact
    # JointDistLoss_mt2: nStack=1
    p1, q = cu(g["mt2_p"])[:, None], cu(g["mt2_q"])[None]
    r = ops.dense_mse(p1, q, coef=cu(g["mt2_w"]).expand(B, J), mask_mode=2, thr=0.8, want_scores=True)
    fin = npy(ops.loss_finalize(r["per_loss"], r["mask"], None))
    np.testing.assert_allclose(fin[0], float(g["mt2_loss"]), rtol=RTOL)
    assert (int(fin[1]), int(fin[2])) == (int(g["mt2_num_pseudo"]), int(g["mt2_num_selected"]))
    np.testing.assert_allclose(npy(r["grad"])[:, 0], g["mt2_grad"], rtol=RTOL, atol=1e-9)
    # JointDistLoss (unmasked, no gate): count = B*J
    r = ops.dense_mse(p1, q)
    fin = npy(ops.loss_finalize(r["per_loss"], None, None))
    np.testing.assert_allclose(fin[0], float(g["dist_loss"]), rtol=RTOL)
    assert int(fin[3]) == int(g["dist_count"])
    np.testing.assert_allclose(npy(r["grad"])[:, 0], g["dist_grad"], rtol=RTOL, atol=1e-9)


def test_dense_mse_odd_shape_and_scale(ops):
    rng = np.random.default_rng(13)
    p = rng.standard_normal((3, 2, 4, 17, 23)).astype(np.float32)
    t = rng.standard_normal((2, 3, 4, 17, 23)).astype(np.float32)
    w = np.array([[1.0], [0.0], [0.5]], np.float32)
    o = O.joint_pseudo3(p, np.stack([t, t], 2), w, 2, 0.5, upstream=0.37)
    gs = torch.tensor([0.37], device="cuda")
    r = ops.dense_mse(cu(p), cu(t), coef=cu(w).expand(3, 4), mask_mode=1, thr=0.5, grad_scale=gs)
    fin = npy(ops.loss_finalize(r["per_loss"], r["mask"], None))
    np.testing.assert_allclose(fin[0], o["loss"], rtol=RTOL)
    np.testing.assert_allclose(npy(r["grad"]), o["grad"], rtol=RTOL, atol=1e-9)
    g2 = r["grad"].clone()
    ops.scale_inplace(g2, torch.tensor([2.0], device="cuda"))
    np.testing.assert_allclose(npy(g2), 2 * npy(r["grad"]), rtol=1e-7)


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
def test_ema_golden_bit_exact(ops):
    g = load("ema")
    for epo in (0, 3, 5000):
        alpha = O.ema_alpha(epo, 0.999)
        params = [cu(g["param%d" % i]) for i in range(5)]
        emas = [cu(g["ema%d" % i]) for i in range(5)]
        plan = ops.EmaPlan(params, emas)
        plan.step(alpha)
        for i in range(5):
            assert np.array_equal(npy(emas[i]), g["epo%d_out%d" % (epo, i)]), (epo, i)
        flat_e, flat_p = cu(g["ema3"]).reshape(-1).clone(), cu(g["param3"]).reshape(-1)
        ops.ema_flat(flat_e, flat_p, alpha)
        assert np.array_equal(npy(flat_e).reshape(33, 17), g["epo%d_out3" % epo])


# ------------------------------------------------------------------------------------------------
# the fused chain
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,select,fuse", [(1, "fixed", True), (1, "fixed", False), (1, "quantile", True),
                                           (2, "quantile", True), (2, "fixed", True)])
def test_pipeline_vs_oracle(ops, M, select, fuse):
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=8, K=4, J=6, M=M, S=2, seed=99 + M, jitter=0.5)
    n = {k: v.numpy() for k, v in d.items()}
    o = O.pseudo_label_chain(n["teacher"], n["student"], n["theta"], n["flip"], n["center"], n["scale"], n["islabeled"],
                             select=select, distThrMax=2.0, lossWeight=0.7)
    cfg = pipeline.StepConfig(select=select, distThrMax=2.0, lossWeight=0.7, fuse_k2=fuse)
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64]).cuda()
    w = pipeline.nega_weights(d["islabeled"].cuda(), 1.0)
    r = pipeline.pseudo_label_step(d["teacher"].cuda(), d["student"].cuda(), d["theta"].cuda(), d["flip"].cuda(), dec, w, cfg)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), o["idx"])
    assert np.array_equal(npy(r["xy"]), o["xy"])
    assert np.array_equal(npy(r["enable"]).astype(bool), o["enable"])              # pseudo-label masks bit-exact
    assert np.array_equal(npy(r["gate"]), o["gate"])
    np.testing.assert_allclose(npy(r["kps"]), o["kps"], rtol=1e-6)
    np.testing.assert_allclose(npy(r["dist"]), o["dist"], rtol=1e-12)
    assert int(r["count"]) == o["count"]
    assert o["enable"].any() and not o["enable"].all()
    tgt = npy(r["target"])
    assert np.array_equal(tgt == 0, o["target"] == 0)
    np.testing.assert_allclose(tgt, o["target"], rtol=RTOL)
    loss = float(r["summary"][0]) * float(r["grad_scale"])
    np.testing.assert_allclose(loss, o["loss"], rtol=RTOL)
    np.testing.assert_allclose(npy(r["grad"]), o["grad"], rtol=RTOL, atol=1e-10)


def test_graphed_step_matches_eager(ops):
    """The CUDA-graph capture of the chain (what bench.py times) gives bit-identical outputs and follows
    new inputs copied into its bound buffers."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=8, K=4, J=6, M=1, S=2, seed=5, jitter=0.5, device="cuda")
    d2 = synth.make_batch(B=8, K=4, J=6, M=1, S=2, seed=6, jitter=0.5, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select="fixed", distThrMax=2.0)
    e = torch.randn(1000, device="cuda"); p = torch.randn(1000, device="cuda")
    plan = ops.EmaPlan([p], [e])
    bufs = {k: d[k].clone() for k in ("teacher", "student", "theta", "flip")}
    g = pipeline.GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], dec, w, cfg, ema=plan, alpha=0.5)
    for data in (d, d2):
        for k in ("teacher", "student", "theta"):
            g.state[k].copy_(data[k])
        g.state["flip"].copy_(data["flip"].to(torch.uint8))
        e_before = e.clone()
        st = g.run()
        ref = pipeline.pseudo_label_step(data["teacher"], data["student"], data["theta"], data["flip"], dec, w, cfg)
        for k in ("idx", "max", "xy", "enable", "gate", "grad", "target", "summary", "grad_scale", "count"):
            assert torch.equal(st[k], ref[k]), k
        assert torch.equal(e, torch.addcmul(e_before * 0.5, p, torch.full_like(p, 0.5)).float()) or torch.allclose(e, 0.5 * e_before + 0.5 * p, rtol=1e-6)


def test_full_size_properties(ops):
    """BASELINE config 2 sizes (B=256, K=8, J=14, 64x64): size-independent properties."""
    from ubpl_b200 import synth, pipeline
    B, K, J = 256, 8, 14
    d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select="fixed", distThrMax=3.0)
    r = pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg)
    # (1) fused decode == decode of the materialised warp, for every map
    for v in (0, K - 1):
        back = ops.warp_materialize(d["teacher"][0, v], d["theta"][v], d["flip"][v])
        mx, ix = back.reshape(B, J, -1).max(-1)
        assert torch.equal(ix.to(torch.int32), r["idx"][0, v])
        assert torch.equal(mx, r["max"][0, v])
        plain = ops.warp_decode(back, None, None, dec)
        assert torch.equal(plain["idx"], r["idx"][0, v]) and torch.equal(plain["xy"], r["xy"][0, v])
    # (2) render -> decode round trip: the peak of a rendered target is the (truncated) key point
    kps = torch.cat([r["kps"].reshape(-1, 2), torch.ones(B * J, 1, device="cuda")], -1)
    hm, kout = ops.render_targets(kps, 64, 64, 256, 256)
    vis = kout[:, 2] > 0
    rt = ops.warp_decode(hm.view(B, J, 64, 64), None, None, None, want_hm=True)
    peak = rt["hm_xy"].reshape(-1, 2) - 1
    centre = torch.trunc(r["kps"].reshape(-1, 2)) / 4
    assert torch.all((peak - centre).abs()[vis] <= 0.5 + 1e-6)
    # (3) gradient is linear in the upstream scale and zero where the gate or the weight is zero
    g = r["grad"]
    dead = (r["gate"] == 0) | (w.view(B, 1) == 0)
    assert torch.all(g[dead[:, None, :].expand(B, 2, J)] == 0)
    assert 0.05 < float(r["enable"].float().mean()) < 0.95
    # (4) loss == sum of per-map losses; count == S * #(gate > 0)
    s = r["summary"]
    assert int(s[3]) * 2 == int(r["count"])
    np.testing.assert_allclose(float(s[0]), float(r["per_loss"].double().sum()), rtol=1e-9)
    # (5) EMA at HG2 scale: idempotent at alpha = 1, equals the student at alpha = 0
    e = torch.randn(8_427_548, device="cuda")
    p = torch.randn(8_427_548, device="cuda")
    e0 = e.clone()
    ops.ema_flat(e, p, 1.0)
    assert torch.equal(e, e0)
    ops.ema_flat(e, p, 0.0)
    assert torch.equal(e, p)
