"""GPU parity tests of the fused / early-release variants added on top of the first CUDA path:
K1 with the K2 epilogue (ubpl_warp_decode_k2) and the early-release staging window,
K3 with the loss reduction in its last CTA (ubpl_render_mse_sum), the one-kernel quantile selector
(ubpl_select_quantile_fused) and the single-graph step with event nodes.  Bars as in test_gpu_parity.py:
indices, masks and float64 dispersions bit-exact; losses within 1e-5 relative."""
import os

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


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


KNOBS = [dict(), dict(UBPL_K1_WARPS=3), dict(UBPL_K1_WARPS=1)]


@pytest.mark.parametrize("knobs", KNOBS)
@pytest.mark.parametrize("name", ["chain_mt", "chain_dual"])
def test_k1_variants_golden(ops, knobs, name):
    """Every staging variant of K1 reproduces the reference's indices, scores and coordinates."""
    g = load(name)
    t = g["teacher"]
    M, K, B, J, H, W = t.shape
    dec = ops.decode_coeffs(torch.as_tensor(g["center"]), torch.as_tensor(g["scale"]), [H, W]).cuda()
    with _Env(**knobs):
        for m in range(M):
            r = ops.warp_decode(cu(t[m]), cu(g["theta"]), cu(g["flip"]), dec)
            assert np.array_equal(npy(r["idx"]).astype(np.int64), g["argmax_idx"][m])
            assert np.array_equal(npy(r["max"]), g["max_val"][m])
            assert np.array_equal(npy(r["xy"]), g["preds_multi"][m])


def _edge_maps():
    H = W = 64
    rng = np.random.default_rng(11)
    maps = np.zeros((4, 3, 10, H, W), np.float32)
    maps[:, :, 0] = 1.0                                    # constant: every tie -> first index
    maps[:, :, 1] = -1.0                                   # all negative
    maps[:, :, 2, 10, 20] = 0.5
    maps[:, :, 2, 40, 3] = 0.5                             # exact tie between two far texels (window must be refused)
    maps[:, :, 3] = rng.standard_normal((H, W)) * 1e-3     # noise only
    maps[:, :, 4, 0, 0] = 2.0                              # peaks in the corners: window hangs over the map edge
    maps[:, :, 5, 63, 63] = 2.0
    maps[:, :, 6] = rng.standard_normal((H, W))
    maps[:, :, 6, 30, 30] = np.nan
    maps[:, :, 7] = np.abs(rng.standard_normal((H, W))) + 5.0
    yy, xx = np.mgrid[0:H, 0:W]
    maps[:, :, 8] = np.exp(-((xx - 3.2) ** 2 + (yy - 60.7) ** 2) / 18.0)           # blob near an edge
    maps[:, :, 9] = np.exp(-((xx - 31.5) ** 2 + (yy - 31.5) ** 2) / 200.0)         # broad plateau: four tied texels
    th = np.zeros((4, 3, 2, 3), np.float32)
    for v in range(4):
        for b in range(3):
            ang = np.deg2rad(rng.uniform(-30, 30))
            sc = rng.uniform(0.7, 1.4)
            th[v, b] = [[np.cos(ang) * sc, -np.sin(ang) * sc, 0], [np.sin(ang) * sc, np.cos(ang) * sc, 0]]
    th[0, 0] = [[1, 0, 0], [0, 1, 0]]
    th[1, 1] = [[0.5, 0, 0.3], [0, 0.5, -0.2]]             # zoom + translation
    th[2, 2] = [[2.5, 0, 0], [0, 2.5, 0]]                  # strong minification: wide footprints leave the window
    fl = (rng.random((4, 3)) < 0.5).astype(np.uint8)
    return maps, th, fl


@pytest.mark.parametrize("knobs", KNOBS)
def test_k1_variants_edge_cases_vs_oracle(ops, knobs):
    maps, th, fl = _edge_maps()
    V, B, J, H, W = maps.shape
    center = np.full((B, 2), 128.0, np.float32)
    scale = np.full((B,), 1.28, np.float32)
    back = np.stack([O.affine_back2(maps[v], th[v], fl[v]) for v in range(V)])
    val, idx = O.argmax_first(back)
    xy = np.stack([O.final_preds(back[v], center, scale, [H, W], "f32") for v in range(V)])
    dec = ops.decode_coeffs(torch.as_tensor(center), torch.as_tensor(scale), [H, W]).cuda()
    with _Env(**knobs):
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        r = ops.warp_decode(cu(maps), cu(th), cu(fl), dec, stats=stats)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val, equal_nan=True)
    assert np.array_equal(npy(r["xy"]), xy)
    assert int(stats[2]) == V * B * J


@pytest.mark.parametrize("knobs", KNOBS[:2])
@pytest.mark.parametrize("mode", [1, 2])
def test_warp_decode_k2_matches_unfused(ops, knobs, mode):
    """The K2 epilogue fused into K1 gives exactly what the separate K2 kernels give on the same decode."""
    from ubpl_b200 import synth
    for (B, K, J, seed) in ((8, 4, 6, 1), (33, 8, 14, 2), (5, 1, 3, 3), (4, 32, 2, 4)):
        d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, seed=seed, jitter=0.7, device="cuda")
        dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
        with _Env(**knobs):
            r = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, mode, S=2, distThrMax=2.0)
            ref = ops.warp_decode(d["teacher"][0], d["theta"], d["flip"], dec)
        for k in ("idx", "max", "xy"):
            assert torch.equal(r[k], ref[k]), k
        if mode == 2:
            k2 = ops.k2_view_fixed(ref["xy"], 2.0, 2, 256, 256, 4.0, 3.0)
            for k in ("mean", "dist", "legal", "enable", "gate"):
                assert torch.equal(r[k], k2[k]), k
            assert torch.equal(r["counts"], k2["counts"])
            assert int(r["count"]) == int(k2["count"])
        else:
            vd = ops.view_dispersion(ref["xy"], sentinel_illegal=True)
            for k in ("mean", "dist", "legal"):
                assert torch.equal(r[k], vd[k]), k
        # the workspace cleans up after itself: arrival counters are back to zero
        assert int(r["ws"][128 + ((J + 3) & ~1):128 + ((J + 3) & ~1) + B * J].abs().sum()) == 0


def test_warp_decode_k2_exhaustive_maps(ops):
    """Maps that need the exhaustive decode (white noise, NaN, constant, singular transforms) through the fused
    entry: the maps decoded cooperatively by the whole CTA hand their coordinates to K2 like any other map."""
    rng = np.random.default_rng(5)
    maps, th, fl = _edge_maps()
    noise = rng.standard_normal((6, 9, 7, 64, 64)).astype(np.float32)          # every map is structure-less
    noise[:, :, 3] = 0.25                                                      # constant: ties everywhere
    noise[2, 4, 5, 17, 40] = np.nan
    noise[1, 2, 6] = -np.abs(noise[1, 2, 6])
    ang = rng.uniform(-0.6, 0.6, (6, 9)); sc = rng.uniform(0.6, 1.5, (6, 9))
    thn = np.zeros((6, 9, 2, 3), np.float32)
    thn[..., 0, 0] = np.cos(ang) * sc; thn[..., 0, 1] = np.sin(ang) * sc
    thn[..., 1, 0] = -np.sin(ang) * sc; thn[..., 1, 1] = np.cos(ang) * sc
    thn[0, 0] = 0.0                                                            # singular
    fln = (rng.random((6, 9)) < 0.5).astype(np.uint8)
    for (m, t, f) in ((maps, th, fl), (noise, thn, fln)):
        K, B, J = m.shape[:3]
        dec = ops.decode_coeffs(torch.full((B, 2), 128.0), torch.full((B,), 1.28), [64, 64]).cuda()
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        r = ops.warp_decode_k2(cu(m), cu(t), cu(f), dec, 2, S=2, distThrMax=2.0, stats=stats)
        ref = ops.warp_decode(cu(m), cu(t), cu(f), dec)
        assert torch.equal(r["idx"], ref["idx"])
        assert np.array_equal(npy(r["max"]), npy(ref["max"]), equal_nan=True)
        assert torch.equal(r["xy"], ref["xy"])
        # ... and both are what the oracle's back-warp + first arg-max give (the cooperative exhaustive decode)
        back = np.stack([O.affine_back2(m[v], t[v], f[v]) for v in range(K)])
        val, idx = O.argmax_first(back)
        assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
        assert np.array_equal(npy(r["max"]), val, equal_nan=True)
        assert int(r["status"]) == 0
        k2 = ops.k2_view_fixed(ref["xy"], 2.0, 2, 256, 256, 4.0, 3.0)
        for k in ("mean", "dist", "legal", "enable", "gate"):
            assert torch.equal(r[k], k2[k]), k
        assert int(r["count"]) == int(k2["count"])
        if m is noise:
            assert int(stats[0]) > 0.5 * K * B * J                             # the cooperative path really was exercised


@pytest.mark.parametrize("mode", [3, 4])
def test_warp_decode_k2_dual_matches_unfused(ops, mode):
    """Two teachers: the assess_pseudo_unc2 epilogue of K1 equals view_dispersion x2 + assess_dual (+ select_fixed +
    gate_prepare) on the same decode, bit for bit."""
    from ubpl_b200 import synth
    for (B, K, J, seed) in ((8, 4, 6, 1), (19, 8, 9, 2), (3, 2, 3, 3), (2, 16, 2, 4), (5, 9, 4, 5)):
        d = synth.make_batch(B=B, K=K, J=J, M=2, S=2, seed=seed, jitter=0.7, device="cuda")
        dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
        t = d["teacher"].reshape(2 * K, B, J, 64, 64)
        th = d["theta"].unsqueeze(0).expand(2, K, B, 2, 3).reshape(2 * K, B, 2, 3)
        fl = d["flip"].unsqueeze(0).expand(2, K, B).reshape(2 * K, B)
        r = ops.warp_decode_k2(t, th, fl, dec, mode, S=2, distThrMax=2.0)
        ref = ops.warp_decode(t, th, fl, dec)
        for k in ("idx", "max", "xy"):
            assert torch.equal(r[k], ref[k]), k
        xy = ref["xy"].view(2, K, B, J, 2)
        vd1, vd2 = ops.view_dispersion(xy[0]), ops.view_dispersion(xy[1])
        ad = ops.assess_dual(vd1["mean"], vd2["mean"], None, xy[0], xy[1])
        assert torch.equal(r["mean"], ad["coord32"])
        assert torch.equal(r["dist"], ad["extDist"])
        assert torch.equal(r["legal"].double(), ad["legal"])
        assert int(r["zero_div"]) == int(ad["zero_div"])
        if mode == 4:
            sel = ops.select_fixed(ad["extDist"], ad["legal"], J, 2.0)
            gate, gs, cnt = ops.gate_prepare(ad["coord32"], sel["gate"], 2, 256, 256, 4.0, 3.0, 1.0)
            assert torch.equal(r["enable"].reshape(-1), sel["enable"])
            assert torch.equal(r["gate"].reshape(-1), gate)
            assert torch.equal(r["counts"], sel["counts"]) and int(r["count"]) == int(cnt)


def test_warp_decode_k2_golden(ops):
    g = load("chain_mt")
    t = g["teacher"]
    M, K, B, J, H, W = t.shape
    dec = ops.decode_coeffs(torch.as_tensor(g["center"]), torch.as_tensor(g["scale"]), [H, W]).cuda()
    r = ops.warp_decode_k2(cu(t[0]), cu(g["theta"]), cu(g["flip"]), dec, 1)
    assert np.array_equal(npy(r["xy"]), g["preds_multi"][0])
    assert np.array_equal(npy(r["mean"]), g["preds_mean"][0])


def test_render_mse_fused_summary(ops):
    g = load("render")
    rng = np.random.default_rng(2)
    for (B, S, J) in ((4, 2, 6), (37, 1, 5), (256, 2, 14)):
        kps = cu(rng.uniform(-10, 266, (B, J, 2)).astype(np.float32))
        gate = cu((rng.random((B, J)) < 0.6).astype(np.float32))
        w = cu((rng.random(B) < 0.7).astype(np.float32))
        pred = cu(rng.standard_normal((B, S, J, 64, 64)).astype(np.float32) * 0.1)
        a = ops.render_mse(kps, gate, w, pred, 256, 256, want_summary=True)
        b = ops.render_mse(kps, gate, w, pred, 256, 256)
        want = ops.loss_finalize(b["per_loss"], None, b["gate_out"])
        assert torch.equal(a["per_loss"], b["per_loss"]) and torch.equal(a["grad"], b["grad"])
        np.testing.assert_allclose(npy(a["summary"]), npy(want), rtol=1e-12)
        assert npy(a["summary"])[1:].tolist() == npy(want)[1:].tolist()          # the counts are exact
        # twice through the same ticket ring: the ticket returned to zero
        a2 = ops.render_mse(kps, gate, w, pred, 256, 256, want_summary=True)
        assert torch.equal(a2["summary"], a["summary"])


def test_select_quantile_fused_vs_oracle(ops):
    rng = np.random.default_rng(31)
    for n, J in ((1, 1), (7, 7), (1000, 10), (4352, 17), (24576, 16), (24577, 17), (100000, 17)):
        dist = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4          # many exact ties
        dist[rng.random(n) < 0.2] = 999.0
        legal = (rng.random(n) < 0.9)
        for pct, rthr in ((0.5, 0.0), (0.1, 0.0), (0.99, 0.0), (0.5, 0.8), (0.0, 0.0), (1.0, 0.0)):
            rel, thr, en = O.filter_dual(dist, legal.astype(np.float64), rthr, pct, 1.0)
            for lg in (cu(legal.astype(np.uint8)), cu(legal.astype(np.float64))):
                s = ops.select_quantile_fused(cu(dist), lg, J, int((n - 1) * pct), rthr, 1.0)
                assert float(s["thr"]) == thr, (n, pct)
                assert np.array_equal(npy(s["enable"]).astype(bool), en)
                assert np.array_equal(npy(s["reliability"]), rel)
                assert int(s["counts"][-1]) == int(en.sum())
    # continuous distances (no ties), all distinct exponents, and the all-illegal corner
    for n in (513, 30000):
        dist = np.minimum(rng.gamma(2.0, 3.0, n) * 10.0 ** rng.integers(-3, 2, n), 900.0)
        legal = np.ones(n)
        rel, thr, en = O.filter_dual(dist, legal, 0.0, 0.37, 1.0)
        s = ops.select_quantile_fused(cu(dist), cu(legal), 3, int((n - 1) * 0.37), 0.0, 1.0)
        assert float(s["thr"]) == thr and np.array_equal(npy(s["enable"]).astype(bool), en)
    dist = np.full(64, 999.0)
    legal = np.zeros(64)
    rel, thr, en = O.filter_dual(dist, legal, 0.0, 0.5, 1.0)
    s = ops.select_quantile_fused(cu(dist), cu(legal), 4, 31, 0.0, 1.0)
    assert float(s["thr"]) == thr and np.array_equal(npy(s["enable"]).astype(bool), en)


def test_select_quantile_fused_gate(ops):
    """The gate_prepare fusion of the selector equals selector + ubpl_gate_prepare."""
    rng = np.random.default_rng(8)
    n, J, S = 4352, 17, 2
    dist = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4
    legal = (rng.random(n) < 0.9).astype(np.uint8)
    kps = cu(rng.uniform(-5, 261, (n, 2)).astype(np.float32))
    a = ops.select_quantile_fused(cu(dist), cu(legal), J, (n - 1) // 2, 0.0, 1.0, gate=(kps, S, 256, 256, 4.0, 3.0, 0.7))
    b = ops.select_quantile_fused(cu(dist), cu(legal), J, (n - 1) // 2, 0.0, 1.0)
    gate, gs, cnt = ops.gate_prepare(kps, b["gate"], S, 256, 256, 4.0, 3.0, 0.7)
    assert torch.equal(a["gate"], gate) and int(a["count"]) == int(cnt) and float(a["grad_scale"]) == float(gs)
    assert torch.equal(a["enable"], b["enable"])


def test_select_quantile_legacy_kernel_still_agrees(ops):
    rng = np.random.default_rng(5)
    n, J = 4352, 17
    dist = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4
    dist[rng.random(n) < 0.2] = 999.0
    legal = (rng.random(n) < 0.9).astype(np.float64)
    a = ops.select_quantile(cu(dist), cu(legal), J, 0.0, 0.5, 1.0)
    with _Env(UBPL_SELECT="legacy"):
        b = ops.select_quantile(cu(dist), cu(legal), J, 0.0, 0.5, 1.0)
    for k in ("thr", "enable", "reliability", "counts"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("select", ["fixed", "quantile"])
@pytest.mark.parametrize("k12", [True, False])
def test_pipeline_fusion_flags_vs_oracle(ops, select, k12):
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=8, K=4, J=6, M=1, S=2, seed=77, jitter=0.5)
    n = {k: v.numpy() for k, v in d.items()}
    o = O.pseudo_label_chain(n["teacher"], n["student"], n["theta"], n["flip"], n["center"], n["scale"], n["islabeled"],
                             select=select, distThrMax=2.0, lossWeight=0.7)
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64]).cuda()
    w = pipeline.nega_weights(d["islabeled"].cuda(), 1.0)
    for fsum in (True, False):
        cfg = pipeline.StepConfig(select=select, distThrMax=2.0, lossWeight=0.7, fuse_k12=k12, fuse_sum=fsum)
        r = pipeline.pseudo_label_step(d["teacher"].cuda(), d["student"].cuda(), d["theta"].cuda(), d["flip"].cuda(), dec, w, cfg)
        assert np.array_equal(npy(r["idx"]).astype(np.int64), o["idx"])
        assert np.array_equal(npy(r["enable"]).astype(bool).reshape(o["enable"].shape), o["enable"])
        assert np.array_equal(npy(r["gate"]).reshape(o["gate"].shape), o["gate"])
        np.testing.assert_allclose(npy(r["dist"]).reshape(o["dist"].shape), o["dist"], rtol=1e-12)
        assert int(r["count"]) == o["count"]
        loss = float(r["summary"][0]) * float(r["grad_scale"])
        np.testing.assert_allclose(loss, o["loss"], rtol=RTOL)
        np.testing.assert_allclose(npy(r["grad"]), o["grad"], rtol=RTOL, atol=1e-10)


@pytest.mark.parametrize("mode", ["single", "stages"])
@pytest.mark.parametrize("select", ["fixed", "quantile"])
def test_graphed_step_modes(ops, mode, select):
    """One graph with event-record nodes (what bench.py times) == per-stage graphs == eager."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=8, K=4, J=6, M=1, S=2, seed=15, jitter=0.5, device="cuda")
    d2 = synth.make_batch(B=8, K=4, J=6, M=1, S=2, seed=16, jitter=0.5, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select=select, distThrMax=2.0)
    e = torch.randn(1000, device="cuda"); p = torch.randn(1000, device="cuda")
    plan = ops.EmaPlan([p], [e])
    bufs = {k: d[k].clone() for k in ("teacher", "student", "theta", "flip")}
    g = pipeline.GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], dec, w, cfg, ema=plan, alpha=0.5,
                             mode=mode)
    assert g.mode == mode
    for data in (d, d2, d):
        for k in ("teacher", "student", "theta"):
            g.state[k].copy_(data[k])
        g.state["flip"].copy_(data["flip"].to(torch.uint8))
        st = g.run()
        torch.cuda.synchronize()
        ref = pipeline.pseudo_label_step(data["teacher"], data["student"], data["theta"], data["flip"], dec, w, cfg)
        for k in ("idx", "max", "xy", "enable", "gate", "grad", "target", "summary", "grad_scale", "count"):
            assert torch.equal(st[k].reshape(-1), ref[k].reshape(-1)), k
        if mode == "single":
            ms = g.stage_ms()
            assert set(ms) >= {"k1", "k2", "k3"} and all(v >= 0.0 for v in ms.values())
            assert 0.0 < sum(ms.values()) < 50.0
    # the EMA forked beside K1's short second launch (event recorded by the C call between its two launches)
    e2 = torch.randn(1000, device="cuda"); p2 = torch.randn(1000, device="cuda")
    g2 = pipeline.GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], dec, w, cfg, ema=ops.EmaPlan([p2], [e2]),
                              alpha=0.5, mode=mode, overlap_ema="slow")
    e_before = e2.clone()
    st2 = g2.run()
    torch.cuda.synchronize()
    for k in ("idx", "enable", "gate", "grad", "summary", "count"):
        assert torch.equal(st2[k].reshape(-1), ref[k].reshape(-1)), k
    assert torch.allclose(e2, 0.5 * e_before + 0.5 * p2, rtol=1e-6)


def test_empty_and_degenerate_shapes(ops):
    """Empty batches and single-view / single-joint inputs through the fused entries."""
    from ubpl_b200 import synth
    dec0 = torch.zeros(0, 4, dtype=torch.float64, device="cuda")
    r = ops.warp_decode_k2(torch.zeros(3, 0, 5, 64, 64, device="cuda"), torch.zeros(3, 0, 2, 3, device="cuda"),
                           torch.zeros(3, 0, dtype=torch.uint8, device="cuda"), dec0, 2)
    assert r["xy"].shape == (3, 0, 5, 2) and int(r["count"]) == 0 and int(r["counts"].sum()) == 0
    d = synth.make_batch(B=3, K=1, J=1, M=1, S=1, seed=2, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    r = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 2, S=1, distThrMax=1.0)
    ref = ops.warp_decode(d["teacher"][0], d["theta"], d["flip"], dec)
    assert torch.equal(r["xy"], ref["xy"]) and torch.all(r["dist"][r["legal"] > 0] == 0)      # one view: no dispersion
    a = ops.render_mse(torch.zeros(0, 4, 2, device="cuda"), None, None, torch.zeros(0, 2, 4, 64, 64, device="cuda"), 256, 256,
                       want_summary=True)
    assert a["summary"].tolist() == [0.0, 0.0, 0.0, 0.0]
    out = ops.view_kps(torch.zeros(0, 4, 3, device="cuda"), torch.zeros(2, 0, 3, 3, dtype=torch.float64, device="cuda"), None, 256)
    assert out.shape == (2, 0, 4, 3)


def test_stream_chunks_matches_per_chunk_steps(ops):
    """pipeline.stream_chunks (c5: a batch streamed through two device buffer sets, H2D of chunk i+1 beside the chain
    of chunk i) gives, chunk by chunk, what pseudo_label_step gives on that slice, and hands every chunk's gradient to
    the callback before its buffers are reused."""
    from ubpl_b200 import synth, pipeline
    B, K, J, chunk = 12, 4, 5, 4
    d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, seed=21, jitter=0.6)
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64]).cuda()
    w = pipeline.nega_weights(d["islabeled"].cuda(), 1.0)
    cfg = pipeline.StepConfig(select="fixed", distThrMax=2.0)
    host = dict(teacher=d["teacher"].pin_memory(), student=d["student"].pin_memory(), theta=d["theta"].pin_memory(),
                flip=d["flip"].to(torch.uint8).pin_memory())
    grads = {}
    out = pipeline.stream_chunks(host, dec, w, cfg, chunk, on_chunk=lambda i, st: grads.__setitem__(i, st["grad"].clone()))
    torch.cuda.synchronize()
    assert len(out) == B // chunk
    for i, b0 in enumerate(range(0, B, chunk)):
        sl = slice(b0, b0 + chunk)
        ref = pipeline.pseudo_label_step(d["teacher"][:, :, sl].cuda(), d["student"][sl].cuda(), d["theta"][:, sl].cuda(),
                                         d["flip"][:, sl].cuda(), dec[sl], w[sl], cfg)
        assert torch.equal(out[i]["summary"], ref["summary"]), i
        assert torch.equal(out[i]["grad_scale"], ref["grad_scale"]) and int(out[i]["count"]) == int(ref["count"])
        assert torch.equal(grads[i], ref["grad"]), i


def test_step_is_repeatable_under_load(ops):
    """Race check by repetition (compute-sanitizer is closed on this pool, profiles/README.md): the c2-size step -- dynamic
    claims, the copy cap, the cooperative exhaustive decode, the fence-free K2 hand-off, the loss reduction in K3's last
    CTA -- replayed 25 times gives bit-identical outputs every time, and the device status word stays clear."""
    from ubpl_b200 import synth, pipeline
    B, K, J = 256, 8, 14
    d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, seed=1388, device="cuda", noise_only_frac=1.0)
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select="fixed", distThrMax=3.0)
    g = pipeline.GraphedStep(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg, instrument=False)
    keys = ("idx", "max", "xy", "kps", "dist", "enable", "gate", "grad", "target", "summary", "count")
    first = None
    for it in range(25):
        st = g.run()
        snap = {k: st[k].clone() for k in keys}
        if first is None:
            first = snap
        else:
            for k in keys:
                assert torch.equal(first[k], snap[k]), (it, k)
    g.check()


@pytest.mark.parametrize("B,K,J", [(8, 4, 6), (64, 8, 14), (0, 4, 6)])
def test_k1_with_ema_in_its_tail(ops, B, K, J):
    """ubpl_warp_decode_k2_ema: the warps that run out of maps do the mean-teacher EMA (utils/parameters.py:4-8).
    K1 / K2 outputs identical to the launch without it; every parameter updated exactly once, bit-identical to the
    oracle's update -- also for tensors smaller than a chunk, of odd length and at unaligned offsets."""
    from ubpl_b200 import synth
    rng = np.random.default_rng(5)
    sizes = [1, 3, 4, 7, 33 * 17, 8192, 8193, 3 * 8192 + 5, 70000]
    base = torch.randn(sum(sizes) + 64, device="cuda")
    params, emas, off = [], [], 1                        # off = 1: the first tensors are not 16-byte aligned
    for n in sizes:
        params.append(base[off:off + n])
        emas.append(cu(rng.standard_normal(n).astype(np.float32)))
        off += n
    plan = ops.EmaPlan(params, emas)
    alpha = O.ema_alpha(3, 0.999)
    want = [O.ema_update(npy(e), npy(p), alpha) for e, p in zip(emas, params)]
    if B == 0:
        t = torch.zeros(K, 0, J, 64, 64, device="cuda")
        th = torch.zeros(K, 0, 2, 3, device="cuda"); fl = torch.zeros(K, 0, dtype=torch.uint8, device="cuda")
        dec = torch.zeros(0, 4, dtype=torch.float64, device="cuda")
        ops.warp_decode_k2(t, th, fl, dec, 2, S=2, distThrMax=2.0, ema=plan, alpha=alpha)
    else:
        d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, seed=31, jitter=0.5, device="cuda")
        dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
        a = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 2, S=2, distThrMax=2.0)
        b = ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 2, S=2, distThrMax=2.0, ema=plan, alpha=alpha,
                               prefetch=d["student"])
        ops.check_status(b["status"])
        for k in ("idx", "max", "xy", "mean", "dist", "legal", "enable", "gate", "counts", "count"):
            assert torch.equal(a[k], b[k]), k
    torch.cuda.synchronize()
    for i, (e, w) in enumerate(zip(emas, want)):
        assert np.array_equal(npy(e), w), (i, sizes[i])
    # alpha from the plan's device buffer (what a captured graph uses), second update on top of the first
    plan.set_alpha(0.25)
    want2 = [O.ema_update(w, npy(p), 0.25) for w, p in zip(want, params)]
    if B > 0:
        ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 1, ema=plan, alpha=0.9, alpha_from_device=True)
        torch.cuda.synchronize()
        for i, (e, w) in enumerate(zip(emas, want2)):
            assert np.array_equal(npy(e), w), (i, sizes[i])


def test_graphed_step_with_ema_in_k1_tail(ops):
    """overlap_ema="tail": the step's EMA rides in K1's launch; same outputs as the eager chain, EMA applied once per replay
    with the alpha of set_alpha()."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=16, K=4, J=6, M=1, S=2, seed=15, jitter=0.5, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    for select in ("fixed", "quantile"):
        cfg = pipeline.StepConfig(select=select, distThrMax=2.0)
        e = torch.randn(50000, device="cuda"); p = torch.randn(50000, device="cuda")
        g = pipeline.GraphedStep(d["teacher"].clone(), d["student"].clone(), d["theta"].clone(), d["flip"].clone(), dec, w, cfg,
                                 ema=ops.EmaPlan([p], [e]), alpha=0.5, overlap_ema="tail")
        assert g.overlap_ema == "tail"
        ref = pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg)
        for alpha in (0.5, 0.9):
            g.set_alpha(alpha)
            before = npy(e).copy()
            st = g.run()
            torch.cuda.synchronize()
            g.check()
            for k in ("idx", "max", "xy", "enable", "gate", "grad", "target", "summary", "grad_scale", "count"):
                assert torch.equal(st[k].reshape(-1), ref[k].reshape(-1)), (select, k)
            assert np.array_equal(npy(e), O.ema_update(before, npy(p), alpha)), (select, alpha)


@pytest.mark.parametrize("M,select", [(1, "fixed"), (2, "quantile")])
def test_eager_step_with_ema(ops, M, select):
    """pseudo_label_step(ema=plan, alpha=a): same chain outputs, the EMA applied once with `a` (by value: no sync)."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=8, K=4, J=6, M=M, S=2, seed=21, jitter=0.5, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select=select, distThrMax=2.0)
    e = torch.randn(30001, device="cuda"); p = torch.randn(30001, device="cuda")
    before = npy(e).copy()
    ref = pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg)
    st = pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg, ema=ops.EmaPlan([p], [e]), alpha=0.8)
    torch.cuda.synchronize()
    for k in ("idx", "max", "xy", "enable", "gate", "grad", "target", "summary", "grad_scale", "count"):
        assert torch.equal(st[k].reshape(-1), ref[k].reshape(-1)), k
    assert np.array_equal(npy(e), O.ema_update(before, npy(p), 0.8))


@pytest.mark.parametrize("hw", [(64, 64), (48, 80)])
def test_k1_border_pixels_vs_oracle(ops, hw):
    """Maps whose warped maximum is not positive and whose frame pokes out of the source map (zero padding at the ends
    of the rows) -- the cases K1 solves geometrically, prunes with the all-inside bound, or decodes exhaustively.
    Non-positive noise, non-positive smooth bumps, an exact-zero plateau, low-contrast negatives and maps sprinkled
    with zeros, under random rotations / scales / shifts; indices, values (the sign of a zero too) and coordinates
    bit-exact against the oracle."""
    H, W = hw
    rng = np.random.default_rng(77)
    V, B, J = 6, 8, 5
    maps = np.empty((V, B, J, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for v in range(V):
        for b in range(B):
            maps[v, b, 0] = -np.abs(rng.standard_normal((H, W))).astype(np.float32) - 1e-3          # negative noise
            cx, cy = rng.uniform(0, W), rng.uniform(0, H)
            maps[v, b, 1] = -1.0 + 0.9 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 18.0)            # negative bump
            maps[v, b, 2] = np.minimum(maps[v, b, 1] + 0.2, 0.0)                                      # zero plateau
            maps[v, b, 3] = (rng.standard_normal((H, W)) * 0.1 - 0.5).astype(np.float32)              # negative, low contrast
            maps[v, b, 4] = -np.abs(rng.standard_normal((H, W))).astype(np.float32) * (rng.random((H, W)) < 0.9)  # zeros sprinkled
    ang = rng.uniform(-0.6, 0.6, (V, B)); sc = rng.uniform(0.75, 1.15, (V, B))
    th = np.zeros((V, B, 2, 3), np.float32)
    th[..., 0, 0] = sc * np.cos(ang); th[..., 0, 1] = -sc * np.sin(ang); th[..., 0, 2] = rng.uniform(-0.2, 0.2, (V, B))
    th[..., 1, 0] = sc * np.sin(ang); th[..., 1, 1] = sc * np.cos(ang); th[..., 1, 2] = rng.uniform(-0.2, 0.2, (V, B))
    fl = (rng.random((V, B)) < 0.5).astype(np.uint8)
    center = np.full((B, 2), 128.0, np.float32)
    scale = np.full((B,), 1.28, np.float32)
    back = np.stack([O.affine_back2(maps[v], th[v], fl[v]) for v in range(V)])
    val, idx = O.argmax_first(back)
    xy = np.stack([O.final_preds(back[v], center, scale, [H, W], "f32") for v in range(V)])
    dec = ops.decode_coeffs(torch.as_tensor(center), torch.as_tensor(scale), [H, W]).cuda()
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = ops.warp_decode(cu(maps), cu(th), cu(fl), dec, stats=stats)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)
    assert np.array_equal(np.signbit(npy(r["max"])), np.signbit(val))         # the sign of a zero maximum too
    assert np.array_equal(npy(r["xy"]), xy)
    assert int(stats[2]) == V * B * J
