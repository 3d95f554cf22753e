"""Pins oracle/ubpl_oracle.py against the UNMODIFIED reference imported from /root/reference.
Runs only where the reference is mounted (the build container); the same comparisons are frozen
into tests/golden/*.npz by tests/golden/make_golden.py for everywhere else."""
import copy
import types

import numpy as np
import pytest
import torch

import ref_import
import ubpl_oracle as O
import ubpl_b200
from ubpl_b200 import synth

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_import.reference_available(), reason="reference not mounted")]


@pytest.fixture(scope="module")
def ref():
    return ref_import.load_reference()


@pytest.fixture(scope="module")
def batch():
    return synth.make_batch(B=6, K=4, J=5, M=2, S=2, seed=1388)


def test_warpmat(ref):
    for ang, sc in [(-20.0, 1 / 1.1), (13.7, 0.8), (0.0, 1.0), (30.0, 1 / 1.6), (-29.9, 1.3)]:
        want = ref.aug.affine_getWarpmat(ang, sc, [256, 256]).numpy()
        got = O.affine_getWarpmat(ang, sc, [256, 256])
        assert np.array_equal(want, got), (ang, sc, want, got)


def test_affine_back2_bit_exact(ref, batch):
    t = batch["teacher"]
    for m in range(t.shape[0]):
        for v in range(t.shape[1]):
            want = ref.aug.affine_back2(t[m, v], batch["theta"][v], batch["flip"][v]).numpy()
            got = O.affine_back2(t[m, v].numpy(), batch["theta"][v].numpy(), batch["flip"][v].numpy())
            assert np.array_equal(want, got)


def test_affine_back2_translation_and_odd_shapes(ref):
    g = torch.Generator().manual_seed(3)
    for (H, W) in [(64, 64), (32, 48), (128, 128), (17, 23)]:
        x = torch.randn(3, 4, H, W, generator=g)
        th = torch.randn(3, 2, 3, generator=g) * 0.6
        fl = torch.tensor([True, False, True])
        want = ref.aug.affine_back2(x, th, fl).numpy()
        got = O.affine_back2(x.numpy(), th.numpy(), fl.numpy())
        assert np.array_equal(want, got), (H, W)


def test_fliplr_back_tensor(ref):
    x = torch.randn(2, 3, 5, 7)
    assert np.array_equal(ref.aug.fliplr_back_tensor(x).numpy(), O.fliplr_back_tensor(x.numpy()))
    assert np.array_equal(ref.aug.fliplr_back_tensor(x[0]).numpy(), O.fliplr_back_tensor(x[0].numpy()))


@pytest.mark.parametrize("scale_kind", ["f32_128", "f32_rand", "f64_rand", "int_one"])
def test_kps_fromHeatmap(ref, batch, scale_kind):
    t = batch["teacher"][0, 0]
    B = t.shape[0]
    g = torch.Generator().manual_seed(5)
    center = batch["center"].clone()
    if scale_kind == "f32_128":
        scale, sd = batch["scale"], "f32"
    elif scale_kind == "f32_rand":
        scale, sd = (0.8 + torch.rand(B, generator=g)).float(), "f32"
        center = center + torch.randint(-20, 20, (B, 2), generator=g)
    elif scale_kind == "f64_rand":
        scale, sd = (0.8 + torch.rand(B, generator=g)).double(), "f64"
        center = (center + torch.randint(-20, 20, (B, 2), generator=g)).double() + 0.5
    else:
        scale, sd = torch.tensor([1 for _ in range(B)]), "f32"
    want_p, want_s = ref.proc.kps_fromHeatmap(t.clone(), center, scale, [64, 64])
    got_p, got_s = O.kps_fromHeatmap(t.numpy(), center.numpy(), scale.numpy(), [64, 64], sd)
    assert np.array_equal(want_p.numpy(), got_p)
    assert np.array_equal(want_s.numpy(), got_s)


def test_kps_fromHeatmap_edge_cases(ref):
    hm = torch.zeros(2, 4, 64, 64)
    hm[0, 0] = 1.0                      # constant map: first index wins
    hm[0, 1] = -1.0                     # all negative: masked -> image coord -3
    hm[0, 2, 10, 20] = 0.5
    hm[0, 2, 40, 3] = 0.5               # tie: row-major first
    hm[1, 0, 63, 63] = 2.0
    hm[1, 1, 0, 0] = 1e-30
    hm[1, 3, 5, 5] = float("nan")
    center = torch.tensor([[128, 128], [128, 128]])
    scale = torch.tensor([1.28, 1.0])
    want_p, want_s = ref.proc.kps_fromHeatmap(hm.clone(), center, scale, [64, 64])
    got_p, got_s = O.kps_fromHeatmap(hm.numpy(), center.numpy(), scale.numpy(), [64, 64])
    assert np.array_equal(want_p.numpy(), got_p)
    assert np.array_equal(want_s.numpy(), got_s, equal_nan=True)


def test_kps_fromHeatmap_mul(ref, batch):
    t = batch["teacher"][0]
    back = torch.stack([ref.aug.affine_back2(t[v], batch["theta"][v], batch["flip"][v]) for v in range(t.shape[0])])
    want = ref.proc.kps_fromHeatmap_mul(back, batch["center"], batch["scale"], [64, 64])
    got = O.kps_fromHeatmap_mul(back.numpy(), batch["center"].numpy(), batch["scale"].numpy(), [64, 64])
    for w, g_ in zip(want, got):
        assert np.array_equal(w.numpy(), g_)


def test_kps_fromHeatmap2(ref, batch):
    s = batch["student"]
    for b in range(3):
        hm = s[b, 0]
        want = ref.proc.kps_fromHeatmap2(hm.clone(), batch["center"][b], batch["scale"][b], [64, 64])
        got = O.kps_fromHeatmap2(hm.numpy(), batch["center"][b].numpy(), batch["scale"][b].numpy(), [64, 64])
        assert np.array_equal(want.numpy(), got)


def _decoded(ref, batch):
    t = batch["teacher"]
    M, K = t.shape[:2]
    out = []
    for m in range(M):
        back = torch.stack([ref.aug.affine_back2(t[m, v], batch["theta"][v], batch["flip"][v]) for v in range(K)])
        out.append(ref.proc.kps_fromHeatmap_mul(back, batch["center"], batch["scale"], [64, 64]))
    return out


def test_uncertainty_fromDistance(ref, batch):
    pm, pbar, _, _ = _decoded(ref, batch)[0]
    want_u, want_w = ref.eval.uncertainty_fromDistance(pm, pbar)
    got_u, got_w = O.uncertainty_fromDistance(pm.numpy(), pbar.numpy())
    assert np.array_equal(want_u.numpy(), got_u)
    np.testing.assert_allclose(want_w.numpy(), got_w, rtol=1e-6)


def _args(**kw):
    base = dict(pck_ref=[0, 1], pck_thr=0.2, br_inferAugNum=4, reliableThr=0.0, reliablePCT=0.5,
                reliableDistMin=1.0, kpsCount=5, distThrMax=1.0)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_assess_and_filter_dual(ref, batch):
    dec = _decoded(ref, batch)
    (pm1, pbar1, _, _), (pm2, pbar2, _, _) = dec
    # "original sample" predictions: use each teacher's view-mean rounded to the integer grid
    p1, p2 = pbar1.round(), pbar2.round()
    # make one joint of both teachers agree exactly across views so intDist hits exact ties
    pmean = ref.bus.preds_mean(p1, p2)
    B, J = p1.shape[:2]
    ids = ["im%d" % b for b in range(B)]
    gt = torch.cat([batch["base_xy"] * 4 + 1, torch.ones(B, J, 1)], -1)
    args = _args(br_inferAugNum=pm1.shape[0], kpsCount=J)
    pseudo, ori_a, aug_a = ref.bus.assess_pseudo_unc2(ids, gt, [p1, p2, pmean], [list(pm1), list(pm2)], args)
    got = O.assess_dual(p1.numpy(), p2.numpy(), pmean.numpy(), pm1.numpy(), pm2.numpy())
    for i, item in enumerate(pseudo):
        b, j = divmod(i, J)
        assert item["coord"] == got["coord"][b, j].tolist()
        assert item["coord_legal"] == got["legal"][b, j]
        for k_ref, k_or in (("intDist1", "intDist1"), ("intDist2", "intDist2"), ("extDist", "extDist"),
                            ("coord_w1", "w1"), ("coord_w2", "w2")):
            assert item[k_ref] == got[k_or][b, j], (i, k_ref)
    for pct in (0.5, 0.25, 0.9):
        args.reliablePCT = pct
        sel, cnt, errs, accs, thr = ref.bus.filter_pseudo2(copy.deepcopy(pseudo), args)
        rel, thr_o, en = O.filter_dual(got["extDist"], got["legal"], args.reliableThr, pct, args.reliableDistMin)
        assert thr == thr_o
        en_ref = {it["kpID"]: it["enable"] for it in sel}
        rel_ref = {it["kpID"]: it["reliability"] for it in sel}
        for i in range(B * J):
            b, j = divmod(i, J)
            kid = "im%d_%d" % (b, j)
            assert en_ref[kid] == int(en[i]) and rel_ref[kid] == rel[i]
        assert cnt[-1] == int(en.sum())


def _mix_inputs(batch, epoch):
    """Inputs of pseudo_cal_unc for one 'epoch': two teachers' predictions, scores, and A = K augmented views per
    key point in the reference's [B,J,A,2] layout; the views drift with the epoch so the LMA history matters."""
    g = torch.Generator().manual_seed(100 + epoch)
    B, J = batch["base_xy"].shape[:2]
    A = 4
    gt = torch.cat([batch["base_xy"] * 4 + 1, torch.ones(B, J, 1)], -1)
    base = (batch["base_xy"] * 4 + 1).round()
    out = [gt]
    for m in range(2):
        preds = base + torch.randint(-1, 2, (B, J, 2), generator=g).float()
        aug = preds[:, :, None, :] + torch.randint(-1 - epoch % 2, 2 + epoch % 2, (B, J, A, 2), generator=g).float()
        aug[0, 0] = preds[0, 0][None]                                   # one key point with identical views: intDist 0
        scores = torch.rand(B, J, generator=g) * 1.2 - 0.1              # some outside [0,1]: _scoreFormat clamps
        aug_scores = torch.rand(B, J, A, generator=g)
        out += [preds, scores, aug, aug_scores]
    return out


def test_pseudo_cal_unc_and_mixUnc_filters(ref, batch):
    """a13: utils/business.py:220-294 over three epochs (the LMA cache carries state between calls)."""
    B, J = batch["base_xy"].shape[:2]
    ids = ["im%d" % b for b in range(B)]
    args = _args(kpsCount=J, distThrMax=2.5, mds1_lma_cache=[], mds2_lma_cache=[])
    h1, h2 = O.new_lma_history(B * J), O.new_lma_history(B * J)
    for epoch in range(4):
        gt, p1, s1, a1, as1, p2, s2, a2, as2 = _mix_inputs(batch, epoch)
        r1, r2 = ref.bus.pseudo_cal_unc(ids, gt, p1, s1, a1, as1, p2, s2, a2, as2, args)
        o1, o2 = O.pseudo_cal_unc(gt[..., :2].numpy(), args.pck_ref, args.pck_thr, args.distThrMax, p1.numpy(), s1.numpy(),
                                  a1.numpy(), as1.numpy(), p2.numpy(), s2.numpy(), a2.numpy(), as2.numpy(), h1, h2)
        for recs, o in ((r1, o1), (r2, o2)):
            assert len(recs) == B * J
            for i, it in enumerate(recs):
                b, j = divmod(i, J)
                assert it["kpID"] == "im%d_%d" % (b, j)
                for k in ("error", "score", "intDist", "extDist", "aExtDist", "intDist_lma", "extDist_lma", "aExtDist_lma",
                          "mixDist", "unc"):
                    assert it[k] == o[k][b, j], (epoch, i, k, it[k], o[k][b, j])
                assert it["acc_flag"] == o["acc_flag"][b, j]
                assert it["coord_aug"] == o["coord_aug"][b, j].tolist()
                for k in ("intDistOK", "intDistOK_lma", "extDistOK", "extDistOK_lma", "aExtDistOK", "aExtDistOK_lma"):
                    assert it[k] == o[k][b, j], (epoch, i, k)
            sel, cnt, errs, accs, thr = ref.bus.pseudo_filter_mixUnc(copy.deepcopy(recs), args)
            f = O.pseudo_filter_mixUnc(o["unc"], o["error"], o["acc_flag"], J, args.distThrMax)
            assert thr == f["uncThr"] and cnt == f["selCounts"] and errs == f["selErrs"] and accs == f["selAccs"]
            assert [it["enable"] for it in sel] == f["enable"].astype(int).tolist()
            sel2, cnt2, errs2, accs2, sthr2, thr2 = ref.bus.pseudo_filter_mixUnc2(copy.deepcopy(recs), args)
            f2 = O.pseudo_filter_mixUnc(o["unc"], o["error"], o["acc_flag"], J, args.distThrMax, score=o["score"])
            assert sthr2 == f2["scoreThr"] and thr2 == f2["uncThr"] and cnt2 == f2["selCounts"]
            assert errs2 == f2["selErrs"] and accs2 == f2["selAccs"]
            assert [it["enable"] for it in sel2] == f2["enable"].astype(int).tolist()
        if epoch >= 1:
            assert (o1["unc"] < 999).any() and (o1["unc"] == 999).any()


def test_decode_coeffs_random_centres(ref):
    """The inverse decode transform for 2000 random (centre, scale) pairs, bit for bit: `python_float / tensor` is
    reciprocal-multiply in torch, which a true division misses by an ulp for a quarter of the pairs."""
    from utils.udaap.transforms import get_transform
    g = torch.Generator().manual_seed(23)
    for dt, kind in ((torch.float32, "f32"), (torch.float64, "f64")):
        centers = torch.randint(60, 200, (2000, 2), generator=g)
        scales = (0.6 + torch.rand(2000, generator=g) * 1.2).to(dt)
        got = O.decode_coeffs(centers.numpy(), scales.numpy(), [64, 64], kind)
        from ubpl_b200 import ops
        prod = ops.decode_coeffs(centers, scales, [64, 64]).numpy()
        for i in range(2000):
            inv = np.linalg.inv(get_transform(centers[i], scales[i], [64, 64], rot=0))
            want = [inv[0, 0], inv[0, 2], inv[1, 1], inv[1, 2]]
            assert got[i].tolist() == want, (kind, i)
            assert prod[i].tolist() == want, (kind, i)


def test_view_matrix_and_view_kps(ref):
    """N1: get_transform with rotation / transform / affine_kps / kps_fliplr (utils/udaap/transforms.py:119-158,
    utils/augment.py:151-156, utils/process.py:239-242), with the float32 0-d tensors affine_mulKps hands over
    and with python floats; integer coordinates bit-exact."""
    from utils.udaap.transforms import get_transform
    g = torch.Generator().manual_seed(17)
    B, J, V, W = 6, 9, 5, 256
    kps = torch.cat([torch.rand(B, J, 2, generator=g) * 250 + 2, torch.ones(B, J, 1)], -1)
    kps[1, 3, 1] = 0.0                                    # invisible key point: left alone (but still mirrored)
    kps[2, 0, :2] = torch.tensor([128.0, 128.0])
    for kind in ("f32", "f64"):
        mats = np.zeros((V, B, 3, 3))
        flips = np.zeros((V, B), bool)
        want = np.zeros((V, B, J, 3), np.float32)
        for v in range(V):
            for b in range(B):
                flip = bool(torch.rand(1, generator=g) < 0.5)
                center = [float(torch.randint(100, 156, (1,), generator=g)), float(torch.randint(100, 156, (1,), generator=g))]
                scale = 1.28 * torch.randn(1, generator=g).mul_(0.25).add_(1).clamp(0.75, 1.25)[0]
                angle = torch.randn(1, generator=g).mul_(30).clamp(-30, 30)[0]
                if v == 0:
                    angle = angle * 0                      # the rot == 0 branch
                if kind == "f64":
                    scale, angle = float(scale), float(angle)
                k = kps[b].clone()
                if flip:
                    k = ref.proc.kps_fliplr(k, W)
                    center[0] = W - center[0]
                want[v, b] = ref.aug.affine_kps(k, center, scale, [W, W], angle).numpy()
                got_t = O.view_matrix(center, scale, [W, W], angle, kind)
                assert np.array_equal(got_t, get_transform(center, scale, [W, W], rot=angle)), (kind, v, b)
                mats[v, b], flips[v, b] = got_t, flip
        got = O.view_kps(kps.numpy(), mats, flips, W)
        assert np.array_equal(got, want), kind


def test_acc_pck(ref, batch):
    """N2: utils/evaluation.py:92-139 (float32 tensors; sums within 1e-6 relative, the summation order of
    torch.sum differs from a sequential one)."""
    g = torch.Generator().manual_seed(5)
    B, J = 9, 7
    gts = torch.cat([torch.rand(B, J, 2, generator=g) * 250, torch.ones(B, J, 1)], -1)
    gts[0, 2, 0] = 0.5                      # invisible key points: gt <= 1
    gts[3, 2, 1] = 1.0
    gts[:, 5, 0] = 0.0                      # a joint that is never valid: accs = -1
    preds = gts[..., :2] + torch.randn(B, J, 2, generator=g) * 12
    for thr in (0.2, 0.5):
        want_e, want_a = ref.eval.acc_pck(preds, gts, [0, 1], thr)
        got_e, got_a = O.acc_pck(preds.numpy(), gts.numpy(), [0, 1], thr)
        np.testing.assert_allclose(got_e, want_e.numpy(), rtol=1e-6)
        np.testing.assert_allclose(got_a, want_a.numpy(), rtol=1e-6)
        assert want_a[5] == -1 and got_a[5] == -1


def test_features_cov(ref):
    """N3: utils/process.py:19-31 and its autograd gradient, 1e-5 relative (float32 matmul / mean orders differ)."""
    g = torch.Generator().manual_seed(9)
    for shape in ((3, 2, 8, 16, 16), (2, 1, 5, 7, 9)):
        a = torch.randn(*shape, generator=g).requires_grad_(True)
        b = (0.3 * a.detach() + torch.randn(*shape, generator=g)).requires_grad_(True)
        val, cnt = ref.proc.features_cov(a, b)
        (val * 1.7).backward()
        v, rows, g1, g2 = O.features_cov(a.detach().numpy(), b.detach().numpy(), upstream=1.7)
        assert rows == cnt
        np.testing.assert_allclose(v, val.item(), rtol=1e-5)
        np.testing.assert_allclose(g1, a.grad.numpy(), rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(g2, b.grad.numpy(), rtol=1e-4, atol=1e-9)


def test_kps_heatmap(ref):
    g = torch.Generator().manual_seed(11)
    kps = torch.rand(40, 3, generator=g) * 270 - 8
    kps[:, 2] = 1
    kps[0, :2] = torch.tensor([3.0, 100.0])     # ul < 0 -> invisible
    kps[1, :2] = torch.tensor([252.0, 100.0])   # br >= w -> invisible
    kps[2, :2] = torch.tensor([251.9, 3.0])
    want_h, want_k = ref.proc.kps_heatmap(kps.clone(), (3, 256, 256), 256, 64)
    got_h, got_k = O.kps_heatmap(kps.numpy(), (3, 256, 256), 256, 64)
    assert np.array_equal(want_h.numpy(), got_h)
    assert np.array_equal(want_k.numpy(), got_k)


def _close(a, b, rtol=1e-5):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=1e-30)


def test_losses(ref, batch):
    p = batch["student"].clone().requires_grad_(True)
    B, S, J = p.shape[:3]
    g = torch.Generator().manual_seed(2)
    tgt = torch.rand(B, J, 64, 64, generator=g)
    gate = (torch.rand(B, J, generator=g) > 0.3).float()
    w = batch["islabeled"].float().unsqueeze(-1)
    crit = ref.losses.JointMSELoss(nStack=S, useKPsGate=True, useSampleWeight=True)
    loss, n = crit(p, tgt, gate, w)
    loss.backward()
    lo, no, go = O.joint_mse(p.detach().numpy(), tgt.numpy(), gate.numpy(), w.numpy(), S, True, True)
    _close(loss.item(), lo)
    assert n == no
    np.testing.assert_allclose(p.grad.numpy(), go, rtol=1e-5, atol=1e-9)
    # kpsGate=None => count = B*J (utils/losses.py:18-19)
    loss2, n2 = ref.losses.JointDistLoss()(p[:, -1], tgt)
    lo2, no2, _ = O.joint_mse(p[:, -1].detach().numpy(), tgt.numpy())
    _close(loss2.item(), lo2)
    assert n2 == no2 == B * J


def test_pseudo_losses(ref, batch):
    p = (batch["student"] * 1.2).clone().requires_grad_(True)
    B, S, J = p.shape[:3]
    t = batch["teacher"][:, 0]                       # [M,B,J,H,W] one view
    targets = torch.stack([t * 0.9, t * 1.2], 2)     # [M,B,S,J,H,W]
    nega = (~batch["islabeled"]).float().unsqueeze(-1)
    crit = ref.losses.JointPseudoLoss3(nStack=S, scoreThr=0.8)
    loss, n_p, n_s, jsm, t1, t2 = crit(p, targets, nega)
    loss.backward()
    o = O.joint_pseudo3(p.detach().numpy(), targets.numpy(), nega.numpy(), S, 0.8)
    _close(loss.item(), o["loss"])
    assert (n_p, n_s) == (o["num_pseudo"], o["num_selected"])
    np.testing.assert_allclose(jsm.detach().numpy(), o["joint_score_mean"], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), o["grad"], rtol=1e-5, atol=1e-9)
    with pytest.raises(RuntimeError):
        crit(p, targets, torch.zeros(B, 1))
    with pytest.raises(RuntimeError):
        O.joint_pseudo3(p.detach().numpy(), targets.numpy(), np.zeros((B, 1), np.float32), S, 0.8)
    # JointDistLoss_mt2 (DualPose_UBPL.py:205)
    q = (t[0] * 1.3)
    p1 = batch["student"][:, -1].clone().requires_grad_(True)
    cons = torch.where(batch["islabeled"], 1.0, 0.7).unsqueeze(-1)
    crit2 = ref.losses.JointDistLoss_mt2(nStack=1, useKPsGate=False, useSampleWeight=True, scoreThr=0.8)
    loss, n, n_p, n_s, jsm = crit2(p1, q, sampleWeight=cons)
    loss.backward()
    o = O.joint_dist_mt2(p1.detach().numpy(), q.numpy(), None, cons.numpy(), 1, False, True, 0.8)
    _close(loss.item(), o["loss"])
    assert (n, n_p, n_s) == (o["count"], o["num_pseudo"], o["num_selected"])
    np.testing.assert_allclose(jsm.detach().numpy(), o["joint_score_mean"], rtol=1e-6)
    np.testing.assert_allclose(p1.grad.numpy(), o["grad"], rtol=1e-5, atol=1e-9)


def test_ema(ref):
    g = torch.Generator().manual_seed(9)
    a = torch.nn.Linear(37, 53)
    b = torch.nn.Linear(37, 53)
    for epo, decay in [(0, 0.999), (3, 0.999), (5000, 0.999)]:
        ema_before = [p.detach().clone().numpy() for p in b.parameters()]
        ref.parameters.update_ema_variables(a, b, types.SimpleNamespace(epo=epo, ema_decay=decay))
        alpha = O.ema_alpha(epo, decay)
        for e0, p, e1 in zip(ema_before, a.parameters(), b.parameters()):
            assert np.array_equal(O.ema_update(e0, p.detach().numpy(), alpha), e1.detach().numpy())


def test_flip_back_swap(ref):
    """utils/udaap/transforms.py:20-57 flip_back (mirror + sequential left/right channel exchange) against the
    oracle's flip_back_swap, and the channel table the CUDA path uses (ops.swap_perm_from_pairs)."""
    from ubpl_b200 import ops
    rng = np.random.default_rng(3)
    for name, J in (("mpii", 16), ("real_animal", 18)):
        x = rng.standard_normal((3, J, 8, 12)).astype(np.float32)
        want = ref.udaap_tf.flip_back(torch.from_numpy(x.copy()), name).numpy()
        got = O.flip_back_swap(x, O.FLIP_PAIRS[name])
        assert np.array_equal(got, want)
        perm = ops.swap_perm_from_pairs(O.FLIP_PAIRS[name], J).numpy()
        assert np.array_equal(x[..., ::-1][:, perm], want)                  # out[:, j] = mirrored[:, perm[j]]


@pytest.mark.parametrize("M,select", [(1, "fixed"), (1, "quantile"), (2, "fixed"), (2, "quantile")])
def test_reference_chain_matches_oracle_chain(ref, M, select):
    """oracle/ref_chain.py (the reference's own functions composed into the benchmark's chain -- the CPU baseline
    bench.py times) against the oracle's restated chain on the same inputs."""
    import ref_chain
    d = synth.make_batch(B=6, K=4, J=5, M=M, S=2, seed=77 + M, jitter=0.7)
    n = {k: v.numpy() for k, v in d.items()}
    o = O.pseudo_label_chain(n["teacher"], n["student"], n["theta"], n["flip"], n["center"], n["scale"], n["islabeled"],
                             select=select, distThrMax=2.0, lossWeight=0.7)
    r = ref_chain.reference_chain(ref, d, select=select, distThrMax=2.0, lossWeight=0.7)
    for k in ("idx", "xy", "enable", "gate"):
        assert np.array_equal(np.asarray(o[k]), np.asarray(r[k])), k
    assert o["count"] == r["count"]
    np.testing.assert_allclose(r["loss"], o["loss"], rtol=1e-6)
    np.testing.assert_allclose(r["grad"], o["grad"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(r["target"], o["target"], rtol=1e-6, atol=1e-8)
