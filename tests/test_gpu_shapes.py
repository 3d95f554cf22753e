"""GPU parity at every BASELINE.json shape class: the fused chain against the oracle at
128x128 (c5's map size), c3 (M = 2, J = 9, K = 8, fixed rule), c4 (J = 17, K = 16, global quantile) -- each on a
batch large enough to exercise the dynamic claims, the cooperative exhaustive decode and the K2 hand-off -- and c2 at
its FULL size against the oracle on a sampled subset of the samples.  Bars: indices, masks, integer coordinates and
float64 dispersions bit-exact; targets, loss and gradient within 1e-5 relative.  Everything goes through the C ABI."""
import numpy as np
import pytest
import torch

import ubpl_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ubpl_b200
    from ubpl_b200 import ops as _ops
    return _ops


def npy(t):
    return t.detach().cpu().numpy()


def _plant_edge_maps(teacher, rng):
    """Overwrites a few teacher maps (in place, CPU tensor [M,K,B,J,H,W]) with the cases that leave the pruned
    decode: white noise, all-negative noise, constant, exact far ties, NaN, +Inf, a corner peak."""
    M, K, B, J, H, W = teacher.shape
    picks = [(rng.integers(M), rng.integers(K), rng.integers(B), rng.integers(J)) for _ in range(14)]
    for n, (m, k, b, j) in enumerate(picks):
        t = teacher[m, k, b, j]
        kind = n % 7
        if kind == 0:
            t.copy_(torch.from_numpy(rng.standard_normal((H, W)).astype(np.float32) * 0.02))
        elif kind == 1:
            t.copy_(torch.from_numpy(-np.abs(rng.standard_normal((H, W))).astype(np.float32) * 0.02 - 1e-3))
        elif kind == 2:
            t.fill_(0.375)
        elif kind == 3:
            t.zero_()
            t[H // 5, W // 3] = 0.5
            t[(4 * H) // 5, (2 * W) // 3] = 0.5
        elif kind == 4:
            t[H // 2, W // 2] = float("nan")
        elif kind == 5:
            t[H // 3, W // 4] = float("inf")
        else:
            t.zero_()
            t[H - 1, W - 1] = 2.0
    return picks


def _run_chain(ops, d, cfg):
    from ubpl_b200 import pipeline
    H, W = d["teacher"].shape[-2:]
    dec = ops.decode_coeffs(d["center"], d["scale"], [H, W]).cuda()
    w = pipeline.nega_weights(d["islabeled"].cuda(), 1.0)
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = pipeline.pseudo_label_step(d["teacher"].cuda(), d["student"].cuda(), d["theta"].cuda(), d["flip"].cuda(), dec, w,
                                   cfg, stats=stats)
    return r, stats.cpu()


def _check_chain(r, o, decode_only_nan_ok=True):
    assert np.array_equal(npy(r["idx"]).astype(np.int64), o["idx"]), "arg-max indices"
    assert np.array_equal(npy(r["max"]), o["max"], equal_nan=True), "scores"
    assert np.array_equal(npy(r["xy"]), o["xy"]), "integer image coordinates"
    assert np.array_equal(npy(r["enable"]).astype(bool), o["enable"]), "pseudo-label masks"
    assert np.array_equal(npy(r["gate"]), o["gate"]), "gates"
    np.testing.assert_allclose(npy(r["dist"]), o["dist"], rtol=1e-12, err_msg="float64 dispersions")
    np.testing.assert_allclose(npy(r["kps"]), o["kps"], rtol=1e-6, err_msg="selected coordinates")
    assert int(r["count"]) == o["count"]
    tgt = npy(r["target"])
    assert np.array_equal(tgt == 0, o["target"] == 0), "target support"
    np.testing.assert_allclose(tgt, o["target"], rtol=RTOL)
    loss = float(r["summary"][0]) * float(r["grad_scale"])
    np.testing.assert_allclose(loss, o["loss"], rtol=RTOL)
    np.testing.assert_allclose(npy(r["grad"]), o["grad"], rtol=RTOL, atol=1e-10)


def _oracle(d, select, **kw):
    n = {k: v.numpy() for k, v in d.items()}
    return O.pseudo_label_chain(n["teacher"], n["student"], n["theta"], n["flip"], n["center"], n["scale"], n["islabeled"],
                                select=select, **kw)


def test_chain_128x128_vs_oracle(ops):
    """c5's map size through the fused chain: 128x128, 3 warps per SM, 64 KB staging buffers; B = 32 so that the
    2048 maps go through the claim counter, the cooperative exhaustive decode (planted maps) and the K2 hand-off."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=32, K=8, J=8, H=128, W=128, M=1, S=2, seed=51, jitter=1.0, noise_only_frac=0.5)
    _plant_edge_maps(d["teacher"], np.random.default_rng(7))
    # synth: a 128x128 map stands for a 512x512 input (centre 256, scale 2.56), stride 4 as in the reference
    o = _oracle(d, "fixed", distThrMax=2.0)
    r, stats = _run_chain(ops, d, pipeline.StepConfig(select="fixed", distThrMax=2.0))
    _check_chain(r, o)
    assert int(stats[2]) == 8 * 32 * 8 and int(stats[0]) >= 6          # the planted maps took the exhaustive path
    assert o["enable"].any() and not o["enable"].all()


def test_decode_128x128_J32_K16_vs_oracle(ops):
    """c5's joint / view counts (J = 32, K = 16) at 128x128 on the fused decode + dispersion (mode 1), B = 4."""
    from ubpl_b200 import synth
    d = synth.make_batch(B=4, K=16, J=32, H=128, W=128, M=1, S=1, seed=52, jitter=1.0)
    _plant_edge_maps(d["teacher"], np.random.default_rng(8))
    n = {k: v.numpy() for k, v in d.items()}
    back = np.stack([O.affine_back2(n["teacher"][0][v], n["theta"][v], n["flip"][v]) for v in range(16)])
    val, idx = O.argmax_first(back)
    xy, kps = O.kps_fromHeatmap_mul(back, n["center"], n["scale"], [128, 128], "f32")[:2]
    dec = ops.decode_coeffs(d["center"], d["scale"], [128, 128]).cuda()
    r = ops.warp_decode_k2(d["teacher"][0].cuda(), d["theta"].cuda(), d["flip"].cuda(), dec, 1)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val, equal_nan=True)
    assert np.array_equal(npy(r["xy"]), xy)
    assert np.array_equal(npy(r["mean"]), kps)
    legal = np.all(xy >= 0, axis=(0, 3))
    want = np.where(legal, O.view_dispersion(xy, kps), 999.0)
    # the distances to the fractional view mean are IEEE sqrt on the device and libm pow in CPython: <= 1 ulp each
    np.testing.assert_allclose(npy(r["dist"]), want, rtol=1e-14)
    assert int(r["status"]) == 0


def test_chain_c3_shape_vs_oracle(ops):
    """c3: DualPose_UBPL, two teachers, J = 9, K = 8, fixed rule (assess_pseudo_unc2 in K1's epilogue), B = 32."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=32, K=8, J=9, M=2, S=2, seed=53, jitter=1.0, noise_only_frac=0.5)
    _plant_edge_maps(d["teacher"], np.random.default_rng(9))
    o = _oracle(d, "fixed", distThrMax=3.0)
    r, stats = _run_chain(ops, d, pipeline.StepConfig(select="fixed", distThrMax=3.0))
    _check_chain(r, o)
    assert int(stats[2]) == 2 * 8 * 32 * 9
    assert o["enable"].any() and not o["enable"].all()


def test_chain_c4_shape_vs_oracle(ops):
    """c4: AP-10K J = 17, K = 16, global-quantile selection (single GPU: the one-kernel selector), B = 32."""
    from ubpl_b200 import synth, pipeline
    d = synth.make_batch(B=32, K=16, J=17, M=1, S=2, seed=54, jitter=1.0, noise_only_frac=0.5)
    _plant_edge_maps(d["teacher"], np.random.default_rng(10))
    o = _oracle(d, "quantile", reliablePCT=0.5, reliableThr=0.0, reliableDistMin=1.0)
    r, stats = _run_chain(ops, d, pipeline.StepConfig(select="quantile", reliablePCT=0.5, reliableThr=0.0,
                                                      reliableDistMin=1.0))
    _check_chain(r, o)
    assert int(stats[2]) == 16 * 32 * 17
    assert o["enable"].any() and not o["enable"].all()


def test_c2_full_size_vs_oracle_subset(ops):
    """BASELINE config 2 at its full size (B = 256, K = 8, J = 14, 28 672 maps) on the GPU; the oracle decodes a
    random subset of the samples.  Per-sample quantities are compared bit for bit; the gradient through the two
    normalisation counts (it is loss_weight / n times a per-sample term, n = global open-gate count)."""
    from ubpl_b200 import synth, pipeline
    B, K, J = 256, 8, 14
    d = synth.make_batch(B=B, K=K, J=J, M=1, S=2, seed=1388)
    r, stats = _run_chain(ops, d, pipeline.StepConfig(select="fixed", distThrMax=3.0))
    assert int(stats[2]) == K * B * J
    sub = np.sort(np.random.default_rng(3).choice(B, 12, replace=False))
    ds = {k: (v[:, :, sub] if k == "teacher" else v[:, sub] if k in ("theta", "flip") else v[sub]) for k, v in d.items()}
    o = _oracle(ds, "fixed", distThrMax=3.0)
    st = torch.as_tensor(sub).cuda()
    assert np.array_equal(npy(r["idx"][:, :, st]).astype(np.int64), o["idx"])
    assert np.array_equal(npy(r["max"][:, :, st]), o["max"])
    assert np.array_equal(npy(r["xy"][:, :, st]), o["xy"])
    assert np.array_equal(npy(r["enable"][st]).astype(bool), o["enable"])
    assert np.array_equal(npy(r["gate"][st]), o["gate"])
    np.testing.assert_allclose(npy(r["dist"][st]), o["dist"], rtol=1e-12)
    tgt = npy(r["target"][st])
    assert np.array_equal(tgt == 0, o["target"] == 0)
    np.testing.assert_allclose(tgt, o["target"], rtol=RTOL)
    n_full, n_sub = int(r["count"]), o["count"]
    assert n_full == 2 * int((npy(r["gate"]) > 0).sum())
    np.testing.assert_allclose(npy(r["grad"][st]) * n_full, o["grad"] * n_sub, rtol=RTOL, atol=1e-7)
    # the loss of the subset from the per-item losses the step keeps
    per = npy(r["per_loss"]).reshape(B, -1)[sub].astype(np.float64).sum()
    np.testing.assert_allclose(per, o["loss_sum"], rtol=RTOL)


@pytest.mark.parametrize("shape", [(3, 5, 16, 64, 64), (2, 4, 18, 32, 48)])
def test_swap_perm_vs_oracle(ops, shape):
    """Optional left/right joint exchange of flip_back (utils/udaap/transforms.py:20-57) folded into the un-flip:
    the materialised warp and the fused decode against the oracle's flip_back_swap on the same maps."""
    V, B, J, H, W = shape
    rng = np.random.default_rng(21)
    maps = rng.standard_normal((V, B, J, H, W)).astype(np.float32) * 0.05
    yy, xx = np.mgrid[0:H, 0:W]
    for v in range(V):
        for b in range(B):
            for j in range(J):
                cx, cy = rng.uniform(6, W - 6), rng.uniform(6, H - 6)
                maps[v, b, j] += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 18.0).astype(np.float32)
    ang = rng.uniform(-0.5, 0.5, (V, B)); sc = rng.uniform(0.7, 1.2, (V, B))
    th = np.zeros((V, B, 2, 3), np.float32)
    th[..., 0, 0] = np.cos(ang) * sc; th[..., 0, 1] = np.sin(ang) * sc
    th[..., 1, 0] = -np.sin(ang) * sc; th[..., 1, 1] = np.cos(ang) * sc
    fl = (rng.random((V, B)) < 0.6).astype(np.uint8)
    pairs = O.FLIP_PAIRS["mpii"] if J == 16 else O.FLIP_PAIRS["real_animal"]
    perm = ops.swap_perm_from_pairs(pairs, J)
    want = np.stack([O.affine_back2_swap(maps[v], th[v], fl[v], pairs) for v in range(V)])
    got = np.stack([npy(ops.warp_materialize(torch.from_numpy(maps[v]).cuda(), torch.from_numpy(th[v]).cuda(),
                                             torch.from_numpy(fl[v]).cuda(), swap_perm=perm)) for v in range(V)])
    assert np.array_equal(got, want)
    val, idx = O.argmax_first(want)
    r = ops.warp_decode(torch.from_numpy(maps).cuda(), torch.from_numpy(th).cuda(), torch.from_numpy(fl).cuda(), None,
                        swap_perm=perm)
    assert np.array_equal(npy(r["idx"]).astype(np.int64), idx)
    assert np.array_equal(npy(r["max"]), val)
    r2 = ops.warp_decode_k2(torch.from_numpy(maps).cuda(), torch.from_numpy(th).cuda(), torch.from_numpy(fl).cuda(), None,
                            1, swap_perm=perm)
    assert torch.equal(r2["idx"], r["idx"]) and torch.equal(r2["xy"], r["xy"])
    # identity table == no table
    r3 = ops.warp_decode(torch.from_numpy(maps).cuda(), torch.from_numpy(th).cuda(), torch.from_numpy(fl).cuda(), None,
                         swap_perm=torch.arange(J))
    r4 = ops.warp_decode(torch.from_numpy(maps).cuda(), torch.from_numpy(th).cuda(), torch.from_numpy(fl).cuda(), None)
    assert torch.equal(r3["idx"], r4["idx"])


@pytest.mark.parametrize("R,n,J,pct", [(2, 192, 6, 0.5), (8, 4352, 17, 0.5), (4, 1000, 9, 0.25), (8, 700, 14, 0.9),
                                        (16, 64, 4, 0.5), (3, 20000, 14, 0.5)])
def test_multirank_selector_emulated(ops, R, n, J, pct):
    """The multi-GPU selector (histograms exchanged through per-rank buffers, flags, gather of the last few keys) with
    its R ranks emulated by one cooperative launch on this GPU: every rank's threshold and masks equal the
    single-GPU selection of the concatenated shards (business.py:173-217 on the whole batch), bit for bit -- with
    ragged shards, ties, illegal items (999) and all-equal distances among the cases."""
    rng = np.random.default_rng(R * 1000 + n)
    npr = [n] * R
    if R == 4:
        npr = [1000, 37, 512, 999]                                     # ragged shards
    d = np.round(rng.gamma(2.0, 2.0, (R, n)) * 4) / 4                   # quarter-pixel grid: many exact ties
    if R == 16:
        d[:] = 1.5                                                      # every key equal
    legal = rng.random((R, n)) > 0.1
    d[~legal] = 999.0
    dist = torch.from_numpy(d).cuda()
    leg = torch.from_numpy(legal.astype(np.uint8)).cuda()
    r = ops.select_quantile_emul(dist, leg, J, 0.0, pct, 1.0, n_per_rank=npr)
    assert int(r["status"].abs().sum()) == 0
    cat_d = torch.cat([dist[i, :npr[i]] for i in range(R)])
    cat_l = torch.cat([leg[i, :npr[i]] for i in range(R)])
    one = ops.select_quantile_fused(cat_d, cat_l, J, int((sum(npr) - 1) * pct), 0.0, 1.0)
    thr = float(one["thr"])
    assert np.array_equal(npy(r["thr"]), np.full(R, thr))
    off = 0
    for i in range(R):
        assert torch.equal(r["enable"][i, :npr[i]], one["enable"][off:off + npr[i]]), i
        assert torch.equal(r["reliability"][i, :npr[i]], one["reliability"][off:off + npr[i]]), i
        off += npr[i]
    # ... and the oracle's filter on the concatenation gives the same threshold and masks
    rel, othr, en = O.filter_dual(npy(cat_d), npy(cat_l).astype(np.float64), 0.0, pct, 1.0)
    assert othr == thr
    assert np.array_equal(npy(one["enable"]).astype(bool), en.reshape(-1))
