"""The reference-facing Python surface (same names/signatures as utils/losses.py, utils/augment.py,
utils/process.py, utils/evaluation.py, utils/business.py, utils/parameters.py) on the GPU, against
the golden vectors produced by the unmodified reference.  These tests read like the reference's
call sites: CPU tensors in / CPU tensors out where the drivers do that, python-int counts, tuple
returns, the same exceptions."""
import json
import os
import types

import numpy as np
import pytest
import torch

import ubpl_oracle as O
from golden_util import GOLDEN, load

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ubpl_b200
    from ubpl_b200 import augment, business, evaluation, losses, parameters, process
    return types.SimpleNamespace(aug=augment.AugmentUtils, bus=business.BusinessUtils, eval=evaluation.EvaluationUtils,
                                 losses=losses, parameters=parameters, proc=process.ProcessUtils)


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x))


def test_losses_modules(pkg):
    g = load("losses")
    p = T(g["student"]).cuda().requires_grad_(True)
    crit = pkg.losses.JointPseudoLoss3(nStack=2, scoreThr=0.8).cuda()
    loss, n_p, n_s, jsm, t1, t2 = crit(p, T(g["targets"]).cuda(), T(g["nega"]).cuda())
    assert isinstance(n_p, int) and isinstance(n_s, int) and (t1, t2) == (0.8, 0.8)
    assert (n_p, n_s) == (int(g["p3_num_pseudo"]), int(g["p3_num_selected"]))
    np.testing.assert_allclose(loss.item(), float(g["p3_loss"]), rtol=RTOL)
    np.testing.assert_allclose(jsm.detach().cpu().numpy(), g["p3_jsm"], rtol=RTOL)
    (0.37 * loss / max(n_p, 1)).backward(retain_graph=True)                 # the drivers scale and divide, MT_UBPL.py:287
    np.testing.assert_allclose(p.grad.cpu().numpy(), 0.37 / max(n_p, 1) * g["p3_grad"], rtol=RTOL, atol=1e-10)
    p.grad = None
    loss.backward()                                                          # second backward through the same node
    np.testing.assert_allclose(p.grad.cpu().numpy(), g["p3_grad"], rtol=RTOL, atol=1e-10)
    with pytest.raises(RuntimeError):                                        # losses.py:201 on an all-labeled batch
        crit(p, T(g["targets"]).cuda(), torch.zeros(p.shape[0], 1, device="cuda"))

    p1 = T(g["mt2_p"]).cuda().requires_grad_(True)
    crit2 = pkg.losses.JointDistLoss_mt2(nStack=1, useKPsGate=False, useSampleWeight=True, scoreThr=0.8)
    loss, n, n_p, n_s, jsm = crit2(p1, T(g["mt2_q"]).cuda(), sampleWeight=T(g["mt2_w"]).cuda())
    assert (n, n_p, n_s) == (int(g["mt2_count"]), int(g["mt2_num_pseudo"]), int(g["mt2_num_selected"]))
    np.testing.assert_allclose(loss.item(), float(g["mt2_loss"]), rtol=RTOL)
    np.testing.assert_allclose(jsm.detach().cpu().numpy(), g["mt2_jsm"], rtol=RTOL)
    loss.backward()
    np.testing.assert_allclose(p1.grad.cpu().numpy(), g["mt2_grad"], rtol=RTOL, atol=1e-10)

    p2 = T(g["mt2_p"]).cuda().requires_grad_(True)
    loss, n = pkg.losses.JointDistLoss()(p2, T(g["mt2_q"]).cuda())          # MT_UBPL.py:251
    assert n == int(g["dist_count"])
    np.testing.assert_allclose(loss.item(), float(g["dist_loss"]), rtol=RTOL)
    loss.backward()
    np.testing.assert_allclose(p2.grad.cpu().numpy(), g["dist_grad"], rtol=RTOL, atol=1e-10)

    c = load("chain_mt")
    ps = T(c["student"]).cuda().requires_grad_(True)
    gts = torch.autograd.Variable(T(c["target"]).cuda(), requires_grad=True)   # tools.py:57-62 wraps targets like this
    crit3 = pkg.losses.JointMSELoss(nStack=2, useKPsGate=True, useSampleWeight=True)
    loss, n = crit3(ps, gts, T(c["gate"]).cuda(), T(c["weight"]).cuda())
    assert n == int(c["mse_count"]) and isinstance(n, int)
    np.testing.assert_allclose(loss.item(), float(c["mse_loss"]), rtol=RTOL)
    loss.backward()
    np.testing.assert_allclose(ps.grad.cpu().numpy(), c["mse_grad"], rtol=RTOL, atol=1e-10)
    # a strided prediction view, as the drivers pass outs[m, a, :, -1] (MT_UBPL.py:251)
    full = torch.randn(4, 3, 3, 32, 32, device="cuda", requires_grad=True)
    tgt = torch.randn(4, 3, 32, 32, device="cuda")
    loss, n = pkg.losses.JointDistLoss()(full[:, -1], tgt)
    want, n_o, g_o = O.joint_mse(full[:, -1].detach().cpu().numpy(), tgt.cpu().numpy())
    np.testing.assert_allclose(loss.item(), want, rtol=RTOL)
    loss.backward()
    np.testing.assert_allclose(full.grad[:, -1].cpu().numpy(), g_o, rtol=RTOL, atol=1e-10)
    assert float(full.grad[:, 0].abs().sum()) == 0.0


def test_augment(pkg):
    g = load("chain_mt")
    for v in range(g["teacher"].shape[1]):
        out = pkg.aug.affine_back2(T(g["teacher"][0, v]), T(g["theta"][v]), T(g["flip"][v]))     # CPU in -> CPU out
        assert out.device.type == "cpu" and np.array_equal(out.numpy(), g["back"][0, v])
        out = pkg.aug.affine_back2(T(g["teacher"][0, v]).cuda(), T(g["theta"][v]).cuda(), T(g["flip"][v]).cuda())
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), g["back"][0, v])
    x = torch.randn(2, 3, 5, 7)
    assert torch.equal(pkg.aug.fliplr_back_tensor(x), x.flip(-1))
    assert torch.equal(pkg.aug.fliplr_back_tensor(x[0].cuda()).cpu(), x[0].flip(-1))
    for ang, sc in [(-20.0, 1 / 1.1), (13.7, 0.8), (0.0, 1.0), (30.0, 1 / 1.6)]:
        assert np.array_equal(pkg.aug.affine_getWarpmat(ang, sc, [256, 256]).numpy(), O.affine_getWarpmat(ang, sc))


def test_process(pkg):
    g = load("decode")
    for k in ("f32_128", "int_one", "f32_rand", "f64_rand"):
        p, s = pkg.proc.kps_fromHeatmap(T(g["hm"]), T(g[k + "_center"]), T(g[k + "_scale"]), [64, 64])   # .cpu() call site
        assert p.device.type == "cpu" and np.array_equal(p.numpy(), g[k + "_preds"]), k
        assert np.array_equal(s.numpy(), g[k + "_scores"], equal_nan=True), k
    one = pkg.proc.kps_fromHeatmap(T(g["hm"][1]), T(g["f32_128_center"][1]), T(g["f32_128_scale"][1]), [64, 64], mode="single")
    assert np.array_equal(one.numpy(), g["f32_128_preds"][1])
    q = pkg.proc.kps_fromHeatmap2(T(g["q_hm"]), torch.tensor([128, 128]), torch.tensor(1.28), [64, 64])
    assert np.array_equal(q.numpy(), g["q_preds"])
    c = load("chain_mt")
    pm, pbar, sm, sbar = pkg.proc.kps_fromHeatmap_mul(T(c["back"][0]), T(c["center"]), T(c["scale"]), [64, 64])
    assert np.array_equal(pm.numpy(), c["preds_multi"][0]) and np.array_equal(pbar.numpy(), c["preds_mean"][0])
    assert np.array_equal(sm.numpy(), c["scores_multi"][0])
    np.testing.assert_allclose(sbar.numpy(), c["scores_mean"][0], rtol=1e-6)
    r = load("render")
    kps = T(r["kps"]).clone()
    hm, k2 = pkg.proc.kps_heatmap(kps, (3, 256, 256), 256, 64)
    assert k2 is kps and np.array_equal(kps.numpy(), r["kps_out"])           # mutates its input like process.py:268
    assert np.array_equal(hm.numpy() == 0, r["heatmap"] == 0)
    np.testing.assert_allclose(hm.numpy(), r["heatmap"], rtol=RTOL)
    assert pkg.proc.kps_getLabeledCount(torch.tensor([[1.0, 0.0], [0.5, -1.0]])) == 2
    assert pkg.proc.coord_distance([0.0, 0.0], [3.0, 4.0]) == 5.0
    with pytest.raises(ZeroDivisionError):
        pkg.proc.coord_avgDistance([[1.0, 2.0]])


def test_evaluation(pkg):
    c = load("chain_mt")
    unc, uncW = pkg.eval.uncertainty_fromDistance(T(c["preds_multi"][0]), T(c["preds_mean"][0]))
    np.testing.assert_allclose(unc.numpy(), c["unc"], rtol=1e-6)
    np.testing.assert_allclose(uncW.numpy(), c["uncW"], rtol=1e-6)


def test_business(pkg):
    g = json.load(open(os.path.join(GOLDEN, "business.json")))
    args = types.SimpleNamespace(**g["args"])
    f = lambda k: torch.tensor(g[k], dtype=torch.float32)
    pseudo, ori_a, aug_a = pkg.bus.assess_pseudo_unc2(g["ids"], f("gt"), [f("p1"), f("p2"), f("pmean")],
                                                      [list(f("pm1")), list(f("pm2"))], args)
    exact = ("kpID", "imageID", "kIdx", "coord", "coord_gt", "coord_legal", "acc_flag", "coord_w1", "coord_w2",
             "intDist1", "intDist2", "extDist")
    assert len(pseudo) == len(g["pseudo"])
    for got, want in zip(pseudo, g["pseudo"]):
        for k in exact:
            assert got[k] == want[k], (want["kpID"], k, got[k], want[k])     # float64 fields bit-exact
        np.testing.assert_allclose(got["error"], want["error"], rtol=1e-12)
    for got_set, want_set in zip(ori_a + [x for a in aug_a for x in a], g["ori_assess"] + [x for a in g["aug_assess"] for x in a]):
        for got, want in zip(got_set, want_set):
            assert got["coord"] == want["coord"] and got["coord_legal"] == want["coord_legal"]
            assert got["acc_flag"] == want["acc_flag"]
            np.testing.assert_allclose(got["error"], want["error"], rtol=1e-12)
    sel, cnt, errs, accs, thr = pkg.bus.filter_pseudo2(pseudo, args)
    assert thr == g["thr"] and cnt == g["counts"]
    assert [it["kpID"] for it in sel] == [it["kpID"] for it in g["sel"]]      # same (stable) order
    assert [it["enable"] for it in sel] == [it["enable"] for it in g["sel"]]   # masks bit-exact
    assert [it["reliability"] for it in sel] == [it["reliability"] for it in g["sel"]]
    np.testing.assert_allclose(errs, g["errs"], rtol=1e-12)
    np.testing.assert_allclose(accs, g["accs"], rtol=1e-12)
    with pytest.raises(IndexError):
        pkg.bus.filter_pseudo2([], args)
    # filter_pseudo (business.py:49-91): distance between the two teachers' own predictions
    import copy
    sel, cnt, errs, accs, thr = pkg.bus.filter_pseudo([copy.deepcopy(ori_a[0]), copy.deepcopy(ori_a[1]), copy.deepcopy(ori_a[2])], args)
    assert thr == g["fp_thr"] and cnt == g["fp_counts"]
    assert [it["kpID"] for it in sel] == [it["kpID"] for it in g["fp_sel"]]
    assert [it["enable"] for it in sel] == [it["enable"] for it in g["fp_sel"]]
    assert [it["reliability"] for it in sel] == [it["reliability"] for it in g["fp_sel"]]
    assert [it["dist"] for it in sel] == [it["dist"] for it in g["fp_sel"]]
    np.testing.assert_allclose(errs, g["fp_errs"], rtol=1e-12) if "fp_errs" in g else None
    # the two teachers agree everywhere: dist_max == dist_min == 0, the reference divides by zero (business.py:63)
    same = [copy.deepcopy(ori_a[0]), copy.deepcopy(ori_a[0]), copy.deepcopy(ori_a[2])]
    with pytest.raises(ZeroDivisionError):
        pkg.bus.filter_pseudo(same, args)


def test_update_ema_variables(pkg):
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.BatchNorm2d(7), torch.nn.Linear(5, 9)).cuda()
    b = torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.BatchNorm2d(7), torch.nn.Linear(5, 9)).cuda()
    for p in b.parameters():
        p.detach_()                                                           # models_ema are built with nograd
    for epo in (0, 3, 2000):
        before = [p.detach().cpu().numpy().copy() for p in b.parameters()]
        bufs = [x.clone() for x in b.buffers()]
        pkg.parameters.update_ema_variables(a, b, types.SimpleNamespace(epo=epo, ema_decay=0.999))
        alpha = O.ema_alpha(epo, 0.999)
        for e0, p, e1 in zip(before, a.parameters(), b.parameters()):
            assert np.array_equal(O.ema_update(e0, p.detach().cpu().numpy(), alpha), e1.detach().cpu().numpy())
        for x, y in zip(bufs, b.buffers()):
            assert torch.equal(x, y)                                          # BN buffers are NOT averaged (parameters.py:7)
    pkg.parameters.update_ema_variables(a, b, 0.99, 10)                       # utils_mt.py:34-39 signature


def test_install_patches_a_module_tree(pkg, monkeypatch):
    """install() rebinds attributes of the reference's modules; exercised on a stand-in tree because the
    reference itself is not mounted on the GPU box."""
    import sys
    names = ["utils", "utils.losses", "utils.augment", "utils.process", "utils.evaluation", "utils.business",
             "utils.parameters", "utils.udaap", "utils.udaap.utils_mt"]
    for n in names:
        monkeypatch.setitem(sys.modules, n, types.ModuleType(n))
    sys.modules["utils.augment"].AugmentUtils = type("AugmentUtils", (), {})
    sys.modules["utils.process"].ProcessUtils = type("ProcessUtils", (), {})
    sys.modules["utils.evaluation"].EvaluationUtils = type("EvaluationUtils", (), {})
    sys.modules["utils.business"].BusinessUtils = type("BusinessUtils", (), {})
    from ubpl_b200 import install
    done = install.install()
    assert sys.modules["utils.losses"].JointPseudoLoss3 is pkg.losses.JointPseudoLoss3
    assert "utils.parameters.update_ema_variables" in done
    x = torch.randn(2, 3, 8, 8)
    out = sys.modules["utils.augment"].AugmentUtils.fliplr_back_tensor(x)
    assert torch.equal(out, x.flip(-1))


def test_business_mixunc(pkg):
    """a13: pseudo_cal_unc / pseudo_filter_mixUnc(2) (utils/business.py:220-294) over four epochs against the
    reference's records (the LMA cache in args carries state), and the device-resident form (ops.mix_dists +
    ops.mix_unc with a MixUncState) against the same records."""
    import copy
    from ubpl_b200 import ops
    g = json.load(open(os.path.join(GOLDEN, "mixunc.json")))
    a = g["args"]
    J = a["kpsCount"]
    gt = torch.tensor(g["gt"], dtype=torch.float32)
    B = gt.shape[0]
    args = types.SimpleNamespace(pck_ref=a["pck_ref"], pck_thr=a["pck_thr"], kpsCount=J, distThrMax=a["distThrMax"],
                                 mds1_lma_cache=[], mds2_lma_cache=[])
    st = [ops.MixUncState(B * J, "cuda"), ops.MixUncState(B * J, "cuda")]
    for ep in g["epochs"]:
        t = {k: torch.tensor(ep[k], dtype=torch.float32) for k in ("p1", "s1", "a1", "as1", "p2", "s2", "a2", "as2")}
        r1, r2 = pkg.bus.pseudo_cal_unc(g["ids"], gt, t["p1"], t["s1"], t["a1"], t["as1"], t["p2"], t["s2"], t["a2"], t["as2"], args)
        d = ops.mix_dists(t["p1"].cuda(), t["s1"].cuda(), t["a1"].cuda(), t["p2"].cuda(), t["s2"].cuda(), t["a2"].cuda(),
                          gt=gt.cuda(), pck_ref=a["pck_ref"], pck_thr=a["pck_thr"])
        for m, (tag, recs) in enumerate((("1", r1), ("2", r2))):
            want = ep["rec" + tag]
            assert len(recs) == len(want)
            for it, w in zip(recs, want):
                assert set(w) <= set(it), set(w) - set(it)
                for k in w:
                    if k in ("error", "aExtDist", "aExtDist_lma", "mixDist", "unc"):
                        # CPython pow on this host vs the host that made the fixture: <= 1 ulp, 999 sentinels exact
                        assert (it[k] == 999.0) == (w[k] == 999.0)
                        np.testing.assert_allclose(it[k], w[k], rtol=1e-14, err_msg=k)
                    else:
                        assert it[k] == w[k], (k, it[k], w[k])
            sel, cnt, errs, accs, thr = pkg.bus.pseudo_filter_mixUnc(copy.deepcopy(recs), args)
            assert [x["enable"] for x in sel] == ep["f" + tag]["enable"] and cnt == ep["f" + tag]["counts"]
            assert thr == ep["f" + tag]["thr"]
            np.testing.assert_allclose(errs, ep["f" + tag]["errs"], rtol=1e-13)
            sel, cnt, errs, accs, sthr, thr = pkg.bus.pseudo_filter_mixUnc2(copy.deepcopy(recs), args)
            assert [x["enable"] for x in sel] == ep["g" + tag]["enable"] and cnt == ep["g" + tag]["counts"]
            assert sthr == ep["g" + tag]["score_thr"]
            # device-resident form of the same epoch
            u = ops.mix_unc(d["int" + tag], d["ext"], d["aext"], J, a["distThrMax"], st[m])
            for k_dev, k_rec in (("int" + tag, "intDist"), ("ext", "extDist"), ("score" + tag, "score")):
                assert d[k_dev].reshape(-1).cpu().tolist() == [w[k_rec] for w in want], k_dev
            assert d["acc" + tag].reshape(-1).cpu().tolist() == [w["acc_flag"] for w in want]
            np.testing.assert_allclose(d["err" + tag].reshape(-1).cpu().numpy(), [w["error"] for w in want], rtol=1e-15)
            np.testing.assert_allclose(d["aext"].reshape(-1).cpu().numpy(), [w["aExtDist"] for w in want], rtol=1e-15)
            assert d["caug" + tag].reshape(-1, 2).cpu().tolist() == [w["coord_aug"] for w in want]
            for k_dev in ("intDist_lma", "extDist_lma"):
                assert u[k_dev].cpu().tolist() == [w[k_dev] for w in want], k_dev
            np.testing.assert_allclose(u["mixDist"].cpu().numpy(), [w["mixDist"] for w in want], rtol=1e-15)
            unc_dev, unc_ref = u["unc"].cpu().numpy(), np.array([w["unc"] for w in want])
            assert np.array_equal(unc_dev == 999.0, unc_ref == 999.0)
            np.testing.assert_allclose(unc_dev, unc_ref, rtol=1e-14)
            assert u["enable"].cpu().tolist() == ep["f" + tag]["enable"]
            assert u["counts"].cpu().tolist() == ep["f" + tag]["counts"]
            # the median-score gate of pseudo_filter_mixUnc2 on a fresh copy of the state (same epoch inputs)
            st2 = ops.MixUncState(B * J, "cuda")
            st2.hist.copy_(st[m].hist); st2.len.copy_(st[m].len)
    # score gate: replay the last epoch on rewound state is not possible (state advanced); check the gate arithmetic alone
    sc = d["score1"]
    thr_t = torch.tensor([ep["g1"]["score_thr"]], dtype=torch.float64, device="cuda")
    stg = ops.MixUncState(B * J, "cuda")
    u1 = ops.mix_unc(d["int1"], d["ext"], d["aext"], J, a["distThrMax"], stg, score=sc, score_thr=thr_t)
    stp = ops.MixUncState(B * J, "cuda")
    u0 = ops.mix_unc(d["int1"], d["ext"], d["aext"], J, a["distThrMax"], stp)
    gated = (sc.reshape(-1) < thr_t)
    assert torch.all(u1["unc"][gated] == 999.0) and torch.equal(u1["unc"][~gated], u0["unc"][~gated])


def test_acc_pck(pkg):
    """N2 (utils/evaluation.py:92-139) against the oracle; float32, 1e-6 relative (summation order)."""
    from ubpl_b200 import ops
    g = torch.Generator().manual_seed(5)
    for (B, J) in ((9, 7), (64, 14), (1, 3)):
        gts = torch.cat([torch.rand(B, J, 2, generator=g) * 250, torch.ones(B, J, 1)], -1)
        gts[0, 2, 0] = 0.5
        gts[:, J - 2, 0] = 0.0                                           # a joint that is never visible: accs = -1
        preds = gts[..., :2] + torch.randn(B, J, 2, generator=g) * 12
        for thr in (0.2, 0.5):
            want_e, want_a = O.acc_pck(preds.numpy(), gts.numpy(), [0, 1], thr)
            errs, accs = pkg.eval.acc_pck(preds, gts, [0, 1], thr)
            assert errs.device.type == "cpu" and errs.dtype == torch.float32 and errs.shape == (J + 1,)
            np.testing.assert_allclose(errs.numpy(), want_e, rtol=1e-6)
            np.testing.assert_allclose(accs.numpy(), want_a, rtol=1e-6)
            assert accs[J - 2] == -1
            e2, a2, d, dr = ops.acc_pck(preds.cuda(), gts.cuda(), [0, 1], thr, want_dists=True)
            assert d.shape == (J, B) and torch.all(d[J - 2] == -1)


def test_features_cov(pkg):
    """N3 (utils/process.py:19-31): value and both gradients against the oracle, 1e-5 relative."""
    g = torch.Generator().manual_seed(9)
    for shape in ((3, 2, 8, 32, 32), (2, 1, 5, 7, 9), (4, 2, 16, 16, 16), (2, 2, 6, 16, 32)):
        a = torch.randn(*shape, generator=g).cuda().requires_grad_(True)
        b = (0.3 * a.detach().cpu() + torch.randn(*shape, generator=g)).cuda().requires_grad_(True)
        val, cnt = pkg.proc.features_cov(a, b)
        (val * 1.7).backward()
        v, rows, g1, g2 = O.features_cov(a.detach().cpu().numpy(), b.detach().cpu().numpy(), upstream=1.7)
        assert cnt == rows and val.dim() == 0
        np.testing.assert_allclose(val.item(), v, rtol=1e-5)
        np.testing.assert_allclose(a.grad.cpu().numpy(), g1, rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(b.grad.cpu().numpy(), g2, rtol=1e-4, atol=1e-9)


def test_view_kps(pkg):
    """N1: key points into the frame of every augmented view (utils/process.py:239-242, utils/augment.py:151-156,
    utils/udaap/transforms.py:151-158): integer coordinates bit-exact against the reference's outputs, then the
    in-frame targets rendered from them equal the targets of the reference's per-view key points."""
    from ubpl_b200 import ops
    g = load("viewkps")
    V, B = g["flips"].shape
    J = g["kps"].shape[1]
    W = int(g["img_w"])
    out = ops.view_kps(T(g["kps"]).cuda(), T(g["mats"]).cuda(), T(g["flips"]).cuda(), W)
    assert np.array_equal(out.cpu().numpy(), g["out"])
    assert np.array_equal(out.cpu().numpy(), O.view_kps(g["kps"], g["mats"], g["flips"], W))
    # single-sample drop-in (no flip inside affine_kps: the Dataset mirrors before, process.py:239-242)
    for v, b in ((1, 0), (3, 4)):
        k = T(g["kps"][b]).clone()
        if g["flips"][v, b]:
            k[:, 0] = W - k[:, 0]
        got = pkg.aug.affine_kps(k, g["centers"][v, b].tolist(), torch.tensor(g["scales"][v, b]), [W, W],
                                 torch.tensor(g["angles"][v, b]))
        assert got.device.type == "cpu" and np.array_equal(got.numpy(), g["out"][v, b])
    from ubpl_b200 import pipeline
    hv, kv = pipeline.view_targets(T(g["kps"][..., :2]).cuda(), T(g["kps"][..., 2]).cuda(), T(g["mats"]).cuda(),
                                   T(g["flips"]).cuda(), 64, 64, W, W)
    assert hv.shape == (V, B, J, 64, 64) and np.array_equal(kv[..., :2].cpu().numpy(), g["out"][..., :2])
    # in-frame rendering: one launch for all V*B*J targets
    hm, kout = ops.render_targets(out.reshape(-1, 3), 64, 64, W, W)
    want_hm, want_k = O.kps_heatmap(g["out"].reshape(-1, 3), (3, W, W), W, 64)
    assert np.array_equal(hm.cpu().numpy() == 0, want_hm == 0)
    np.testing.assert_allclose(hm.cpu().numpy(), want_hm, rtol=RTOL)
    assert np.array_equal(kout.cpu().numpy(), want_k)


def test_grouped_criteria(pkg):
    """N4: all M*A criterion calls of a driver step in one kernel and one sync == the loop of separate calls."""
    g = torch.Generator().manual_seed(3)
    G, B, S, J = 6, 5, 2, 4
    preds = (torch.rand(G, B, S, J, 32, 32, generator=g) * 0.5).cuda().requires_grad_(True)
    tg_mse = torch.rand(G, B, J, 32, 32, generator=g).cuda()
    tg_dist = (torch.rand(G, B, S, J, 32, 32, generator=g) * 0.5).cuda()
    gate = (torch.rand(G, B, J, generator=g) < 0.7).float().cuda()
    w = (torch.rand(B, 1, generator=g) < 0.6).float().cuda()
    wts = torch.rand(G, generator=g).cuda()
    for crit, tg in ((pkg.losses.JointMSELoss(nStack=S, useKPsGate=True, useSampleWeight=True), tg_mse),
                     (pkg.losses.JointDistLoss(nStack=S, useKPsGate=False, useSampleWeight=True), tg_dist)):
        sums, counts = pkg.losses.grouped(crit, preds, tg, gate, w)
        (sums * wts).sum().backward()
        g_batched = preds.grad.clone()
        preds.grad = None
        want_s, want_c = [], []
        for k in range(G):
            l, n = crit(preds[k], tg[k], gate[k], w)
            want_s.append(l)
            want_c.append(n)
        (torch.stack(want_s) * wts).sum().backward()
        assert counts == want_c
        np.testing.assert_allclose(sums.detach().cpu().numpy(), torch.stack(want_s).detach().cpu().numpy(), rtol=1e-6)
        np.testing.assert_allclose(g_batched.cpu().numpy(), preds.grad.cpu().numpy(), rtol=1e-6, atol=1e-12)
        preds.grad = None


def test_grouped_pseudo_loss3(pkg):
    """N4: the M*K JointPseudoLoss3 calls of the epc loop (projects/MT_UBPL.py:270-298) in one kernel and one sync ==
    the loop of separate calls (sums, python-int counts, joint scores, gradients), with the teacher stacks addressed
    in place through strides (outs_ema[:, a] of every view, last stack)."""
    g = torch.Generator().manual_seed(4)
    M, K, B, S, J, Mt = 2, 3, 5, 2, 4, 2
    blob = torch.zeros(1, K, B, 1, J, 32, 32)
    cx = torch.randint(6, 26, (K, B, J), generator=g)
    cy = torch.randint(6, 26, (K, B, J), generator=g)
    ys, xs = torch.arange(32).view(32, 1).float(), torch.arange(32).view(1, 32).float()
    for k in range(K):
        for b in range(B):
            for j in range(J):
                blob[0, k, b, 0, j] = torch.exp(-((xs - cx[k, b, j]) ** 2 + (ys - cy[k, b, j]) ** 2) / 18.0)
    outs = ((0.8 + 0.3 * torch.rand(M, K, B, S, J, 1, 1, generator=g)) * blob).cuda().requires_grad_(True)
    outs_ema = ((0.8 + 0.3 * torch.rand(Mt, K, B, S, J, 1, 1, generator=g)) * blob).cuda()
    w = torch.tensor([[1.0], [1.0], [0.0], [1.0], [0.0]]).cuda()
    crit = pkg.losses.JointPseudoLoss3(nStack=S, scoreThr=0.95)
    wts = torch.rand(M * K, generator=g).cuda()
    preds = outs.reshape(M * K, B, S, J, 32, 32)
    # group g = m*K + a is compared with the teachers' view a: [Mt, G, B, S, J, H, W] without a copy
    tg = outs_ema.unsqueeze(1).expand(Mt, M, K, B, S, J, 32, 32).reshape(Mt, M * K, B, S, J, 32, 32)
    sums, n_pseudo, n_sel, jsm, t1, t2 = pkg.losses.grouped(crit, preds, tg, sampleWeight=w)
    (sums * wts).sum().backward()
    g_batched = outs.grad.clone()
    outs.grad = None
    want = [crit(outs[m, a], outs_ema.clone()[:, a].detach(), w) for m in range(M) for a in range(K)]
    (torch.stack([x[0] for x in want]) * wts).sum().backward()
    assert n_pseudo == [x[1] for x in want] and n_sel == [x[2] for x in want]
    assert sum(n_sel) > 0
    np.testing.assert_allclose(sums.detach().cpu().numpy(), torch.stack([x[0] for x in want]).detach().cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(jsm.cpu().numpy(), torch.stack([x[3] for x in want]).cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(g_batched.cpu().numpy(), outs.grad.cpu().numpy(), rtol=1e-6, atol=1e-12)
    assert t1 == t2 == 0.95
