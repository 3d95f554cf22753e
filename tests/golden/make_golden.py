"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_import.py) on seeded synthetic inputs.  Run in the build
container only:   python tests/golden/make_golden.py
The fixtures carry inputs AND the reference's outputs so that tests/test_oracle_golden.py and the
GPU parity tests can check against the reference where it is not mounted (the GPU box)."""
import copy
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_import  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import synth  # noqa: E402

ref = ref_import.load_reference()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def chain_fixture(name, B, K, J, H, W, M, seed):
    d = synth.make_batch(B=B, K=K, J=J, H=H, W=W, M=M, S=2, seed=seed)
    t = d["teacher"]
    res = [H, W]
    back = torch.stack([torch.stack([ref.aug.affine_back2(t[m, v], d["theta"][v], d["flip"][v])
                                     for v in range(K)]) for m in range(M)])
    maxv, idx = torch.max(back.reshape(M, K, B, J, -1), -1)
    dec = [ref.proc.kps_fromHeatmap_mul(back[m], d["center"], d["scale"], res) for m in range(M)]
    unc, uncW = ref.eval.uncertainty_fromDistance(dec[0][0], dec[0][1])
    arrs = dict(teacher=t, student=d["student"], theta=d["theta"], flip=d["flip"], center=d["center"],
                scale=d["scale"], islabeled=d["islabeled"], back=back, argmax_idx=idx, max_val=maxv,
                preds_multi=torch.stack([x[0] for x in dec]), preds_mean=torch.stack([x[1] for x in dec]),
                scores_multi=torch.stack([x[2] for x in dec]), scores_mean=torch.stack([x[3] for x in dec]),
                unc=unc, uncW=uncW)
    if M == 2:
        pm1, pm2 = dec[0][0], dec[1][0]
        p1, p2 = dec[0][1].round(), dec[1][1].round()
        pmean = ref.bus.preds_mean(p1, p2)
        ids = ["im%d" % b for b in range(B)]
        gt = torch.cat([d["base_xy"] * 4 + 1, torch.ones(B, J, 1)], -1)
        args = types.SimpleNamespace(pck_ref=[0, 1], pck_thr=0.2, br_inferAugNum=K, reliableThr=0.0,
                                     reliablePCT=0.5, reliableDistMin=1.0, kpsCount=J)
        pseudo, _, _ = ref.bus.assess_pseudo_unc2(ids, gt, [p1, p2, pmean], [list(pm1), list(pm2)], args)
        keys = ["coord_legal", "intDist1", "intDist2", "extDist", "coord_w1", "coord_w2", "error", "acc_flag"]
        for k in keys:
            arrs["dual_" + k] = np.array([it[k] for it in pseudo], np.float64).reshape(B, J)
        arrs["dual_coord"] = np.array([it["coord"] for it in pseudo], np.float64).reshape(B, J, 2)
        arrs["dual_p1"], arrs["dual_p2"], arrs["dual_pmean"], arrs["dual_gt"] = p1, p2, pmean, gt
        for pct in (0.25, 0.5, 0.9):
            args.reliablePCT = pct
            sel, cnt, errs, accs, thr = ref.bus.filter_pseudo2(copy.deepcopy(pseudo), args)
            order = {it["kpID"]: it for it in sel}
            en = np.array([[order["im%d_%d" % (b, j)]["enable"] for j in range(J)] for b in range(B)], np.int32)
            rel = np.array([[order["im%d_%d" % (b, j)]["reliability"] for j in range(J)] for b in range(B)], np.float64)
            tag = "filt%02d_" % int(pct * 100)
            arrs[tag + "enable"], arrs[tag + "reliability"], arrs[tag + "thr"] = en, rel, np.float64(thr)
            arrs[tag + "counts"] = np.array(cnt, np.int64)
    # Gaussian targets at the decoded view-mean coordinates (process.py:253-278) + JointMSELoss on them
    pm = dec[0][1]
    kps = torch.cat([pm, torch.ones(B, J, 1)], -1)
    inp = H * 4
    rend = [ref.proc.kps_heatmap(kps[b].clone(), (3, inp, inp), inp, H) for b in range(B)]
    target = torch.stack([r[0] for r in rend])
    gate = torch.stack([r[1][:, 2] for r in rend])
    w = (~d["islabeled"]).float().unsqueeze(-1)
    p = d["student"].clone().requires_grad_(True)
    loss, n = ref.losses.JointMSELoss(nStack=2, useKPsGate=True, useSampleWeight=True)(p, target, gate, w)
    loss.backward()
    arrs.update(target=target, gate=gate, weight=w, mse_loss=loss, mse_count=np.int64(n), mse_grad=p.grad)
    if M == 2:           # keep the fixture small: the dual fixture pins idx/max/coords, not the maps
        for k in ("back", "mse_grad", "student", "target"):
            arrs.pop(k)
    save(name, **arrs)


def decode_fixture():
    g = torch.Generator().manual_seed(21)
    hm = torch.rand(4, 4, 64, 64, generator=g) - 0.2
    hm[0, 0] = 1.0
    hm[0, 1] = -1.0
    hm[1, 2] = 0.0
    hm[1, 2, 10, 20] = 0.5
    hm[1, 2, 40, 3] = 0.5
    hm[2, 0] = 0.0
    hm[2, 0, 63, 63] = 2.0
    hm[3, 3, 5, 5] = float("nan")
    out = dict(hm=hm)
    cases = {
        "f32_128": (torch.full((4, 2), 128), torch.full((4,), 1.28)),
        "f32_one": (torch.full((4, 2), 128), torch.ones(4)),
        "int_one": (torch.full((4, 2), 128), torch.tensor([1, 1, 1, 1])),
        "f32_rand": (torch.randint(100, 156, (4, 2), generator=g), (0.8 + torch.rand(4, generator=g)).float()),
        "f64_rand": (torch.randint(100, 156, (4, 2), generator=g).double() + 0.5, (0.8 + torch.rand(4, generator=g)).double()),
    }
    for k, (c, s) in cases.items():
        p, sc = ref.proc.kps_fromHeatmap(hm.clone(), c, s, [64, 64])
        out[k + "_center"], out[k + "_scale"], out[k + "_preds"], out[k + "_scores"] = c, s, p, sc
    # quarter-offset decoder (process.py:345-379), single image
    hm2 = torch.rand(5, 64, 64, generator=g)
    out["q_hm"] = hm2
    out["q_preds"] = ref.proc.kps_fromHeatmap2(hm2.clone(), torch.tensor([128, 128]), torch.tensor(1.28), [64, 64])
    save("decode", **out)


def render_fixture():
    g = torch.Generator().manual_seed(11)
    kps = torch.rand(12, 3, generator=g) * 270 - 8
    kps[:, 2] = 1
    kps[0, :2] = torch.tensor([3.0, 100.0])
    kps[1, :2] = torch.tensor([252.0, 100.0])
    kps[2, :2] = torch.tensor([251.9, 3.0])
    kps[3, :2] = torch.tensor([128.0, 128.0])
    kps[4, :2] = torch.tensor([130.0, 131.0])
    h, k2 = ref.proc.kps_heatmap(kps.clone(), (3, 256, 256), 256, 64)
    save("render", kps=kps, heatmap=h, kps_out=k2)


def loss_fixture():
    d = synth.make_batch(B=4, K=2, J=3, H=32, W=32, M=2, S=2, seed=77)
    p = (d["student"] * 1.2).clone().requires_grad_(True)
    t = d["teacher"][:, 0]
    targets = torch.stack([t * 0.9, t * 1.2], 2)
    nega = (~d["islabeled"]).float().unsqueeze(-1)
    crit = ref.losses.JointPseudoLoss3(nStack=2, scoreThr=0.8)
    loss, n_p, n_s, jsm, _, _ = crit(p, targets, nega)
    loss.backward()
    out = dict(student=p.detach(), targets=targets, nega=nega, p3_loss=loss, p3_num_pseudo=np.int64(n_p),
               p3_num_selected=np.int64(n_s), p3_jsm=jsm, p3_grad=p.grad)
    q = t[0] * 1.3
    p1 = d["student"][:, -1].clone().requires_grad_(True)
    cons = torch.where(d["islabeled"], 1.0, 0.7).unsqueeze(-1)
    loss, n, n_p, n_s, jsm = ref.losses.JointDistLoss_mt2(nStack=1, useSampleWeight=True, scoreThr=0.8)(p1, q, sampleWeight=cons)
    loss.backward()
    out.update(mt2_p=p1.detach(), mt2_q=q, mt2_w=cons, mt2_loss=loss, mt2_count=np.int64(n), mt2_num_pseudo=np.int64(n_p),
               mt2_num_selected=np.int64(n_s), mt2_jsm=jsm, mt2_grad=p1.grad)
    p2 = d["student"][:, -1].clone().requires_grad_(True)
    loss, n = ref.losses.JointDistLoss()(p2, q)
    loss.backward()
    out.update(dist_loss=loss, dist_count=np.int64(n), dist_grad=p2.grad)
    save("losses", **out)


def ema_fixture():
    g = torch.Generator().manual_seed(9)
    shapes = [(7, 3, 3, 3), (7,), (1,), (33, 17), (257,)]
    params = [torch.randn(*s, generator=g) * 0.02 for s in shapes]
    emas = [torch.randn(*s, generator=g) * 0.02 for s in shapes]
    out = {}
    for epo in (0, 3, 5000):
        a = types.SimpleNamespace(parameters=lambda: [torch.nn.Parameter(p.clone()) for p in params])
        eparams = [torch.nn.Parameter(e.clone()) for e in emas]
        b = types.SimpleNamespace(parameters=lambda: eparams)
        ref.parameters.update_ema_variables(a, b, types.SimpleNamespace(epo=epo, ema_decay=0.999))
        for i, e in enumerate(eparams):
            out["epo%d_out%d" % (epo, i)] = e.detach()
    for i, (p, e) in enumerate(zip(params, emas)):
        out["param%d" % i], out["ema%d" % i] = p, e
    save("ema", **out)


def business_fixture():
    """assess_pseudo_unc2 + filter_pseudo2 records (utils/business.py:109-217) as JSON."""
    import json
    d = synth.make_batch(B=5, K=3, J=4, H=32, W=32, M=2, S=1, seed=31, jitter=0.7)
    t = d["teacher"]
    M, K, B, J = t.shape[:4]
    dec = []
    for m in range(M):
        back = torch.stack([ref.aug.affine_back2(t[m, v], d["theta"][v], d["flip"][v]) for v in range(K)])
        dec.append(ref.proc.kps_fromHeatmap_mul(back, d["center"], d["scale"], [32, 32]))
    pm1, pm2 = dec[0][0], dec[1][0]
    p1, p2 = dec[0][1].round(), dec[1][1].round()
    pmean = ref.bus.preds_mean(p1, p2)
    ids = ["img_%d" % b for b in range(B)]
    gt = torch.cat([d["base_xy"] * 4 + 1, torch.ones(B, J, 1)], -1)
    args = types.SimpleNamespace(pck_ref=[0, 1], pck_thr=0.5, br_inferAugNum=K, reliableThr=0.0, reliablePCT=0.5,
                                 reliableDistMin=1.0, kpsCount=J)
    pseudo, ori_a, aug_a = ref.bus.assess_pseudo_unc2(ids, gt, [p1, p2, pmean], [list(pm1), list(pm2)], args)
    sel, cnt, errs, accs, thr = ref.bus.filter_pseudo2(copy.deepcopy(pseudo), args)
    fp_sel, fp_cnt, fp_errs, fp_accs, fp_thr = ref.bus.filter_pseudo([copy.deepcopy(ori_a[0]), copy.deepcopy(ori_a[1]), copy.deepcopy(ori_a[2])], args)
    out = dict(fp_sel=fp_sel, fp_counts=fp_cnt, fp_errs=[float(e) for e in fp_errs], fp_accs=[float(a) for a in fp_accs], fp_thr=fp_thr,
               ids=ids, gt=gt.tolist(), p1=p1.tolist(), p2=p2.tolist(), pmean=pmean.tolist(), pm1=pm1.tolist(),
               pm2=pm2.tolist(), args=vars(args), pseudo=pseudo, ori_assess=ori_a, aug_assess=aug_a, sel=sel,
               counts=cnt, errs=[float(e) for e in errs], accs=[float(a) for a in accs], thr=thr)
    json.dump(out, open(os.path.join(HERE, "business.json"), "w"))
    print("business", len(pseudo), "records", sum(cnt[:-1]), "selected")


def mixunc_fixture():
    """a13 (business.py:220-294): four successive calls of pseudo_cal_unc on drifting predictions (the LMA cache
    carries state), each followed by both filters.  Inputs and the reference's records / filter results."""
    B, J, A = 5, 6, 4
    ids = ["im%d" % b for b in range(B)]
    g = torch.Generator().manual_seed(2024)
    gt = torch.cat([torch.randint(20, 236, (B, J, 2), generator=g).float() + 0.5 * torch.randint(0, 2, (B, J, 2), generator=g),
                    torch.ones(B, J, 1)], -1)
    args = types.SimpleNamespace(pck_ref=[0, 1], pck_thr=0.2, kpsCount=J, distThrMax=2.5, mds1_lma_cache=[], mds2_lma_cache=[])
    epochs = []
    for epoch in range(4):
        inp = {}
        for m in ("1", "2"):
            preds = gt[..., :2].round() + torch.randint(-1, 2, (B, J, 2), generator=g).float()
            aug = preds[:, :, None, :] + torch.randint(-1 - epoch % 2, 2 + epoch % 2, (B, J, A, 2), generator=g).float()
            aug[0, 0] = preds[0, 0][None]
            inp["p" + m], inp["a" + m] = preds, aug
            inp["s" + m] = torch.rand(B, J, generator=g) * 1.2 - 0.1
            inp["as" + m] = torch.rand(B, J, A, generator=g)
        r1, r2 = ref.bus.pseudo_cal_unc(ids, gt, inp["p1"], inp["s1"], inp["a1"], inp["as1"], inp["p2"], inp["s2"], inp["a2"],
                                        inp["as2"], args)
        ep = {k: v.tolist() for k, v in inp.items()}
        ep["rec1"], ep["rec2"] = copy.deepcopy(r1), copy.deepcopy(r2)
        for tag, recs in (("1", r1), ("2", r2)):
            sel, cnt, errs, accs, thr = ref.bus.pseudo_filter_mixUnc(copy.deepcopy(recs), args)
            ep["f" + tag] = dict(enable=[it["enable"] for it in sel], counts=cnt, errs=[float(e) for e in errs],
                                 accs=[float(a) for a in accs], thr=thr)
            sel, cnt, errs, accs, sthr, thr = ref.bus.pseudo_filter_mixUnc2(copy.deepcopy(recs), args)
            ep["g" + tag] = dict(enable=[it["enable"] for it in sel], counts=cnt, errs=[float(e) for e in errs],
                                 accs=[float(a) for a in accs], thr=thr, score_thr=sthr, unc=[it["unc"] for it in sel])
        epochs.append(ep)
    out = dict(ids=ids, gt=gt.tolist(), args=dict(pck_ref=[0, 1], pck_thr=0.2, kpsCount=J, distThrMax=2.5), epochs=epochs)
    json.dump(out, open(os.path.join(HERE, "mixunc.json"), "w"))
    print("mixunc", len(epochs), "epochs,", sum(e["f1"]["counts"][-1] for e in epochs), "selections")


def viewkps_fixture():
    """N1: canonical key points into the frames of V augmented views with the reference's own functions
    (kps_fliplr, get_transform with the float32 0-d tensors of affine_mulKps, affine_kps)."""
    from utils.udaap.transforms import get_transform
    g = torch.Generator().manual_seed(17)
    B, J, V, W = 6, 9, 5, 256
    kps = torch.cat([torch.rand(B, J, 2, generator=g) * 250 + 2, torch.ones(B, J, 1)], -1)
    kps[1, 3, 1] = 0.0
    kps[2, 0, :2] = torch.tensor([128.0, 128.0])
    kps[:, :, :2] = (kps[:, :, :2] * 4).round() / 4
    centers = np.zeros((V, B, 2)); scales = np.zeros((V, B), np.float32); angles = np.zeros((V, B), np.float32)
    flips = np.zeros((V, B), np.uint8); mats = np.zeros((V, B, 3, 3)); out = np.zeros((V, B, J, 3), np.float32)
    for v in range(V):
        for b in range(B):
            flip = bool(torch.rand(1, generator=g) < 0.5)
            center = [float(torch.randint(100, 156, (1,), generator=g)), float(torch.randint(100, 156, (1,), generator=g))]
            scale = 1.28 * torch.randn(1, generator=g).mul_(0.25).add_(1).clamp(0.75, 1.25)[0]
            angle = torch.randn(1, generator=g).mul_(30).clamp(-30, 30)[0]
            if v == 0:
                angle = angle * 0
            k = kps[b].clone()
            if flip:
                k = ref.proc.kps_fliplr(k, W)
                center[0] = W - center[0]
            out[v, b] = ref.aug.affine_kps(k, center, scale, [W, W], angle).numpy()
            mats[v, b] = get_transform(center, scale, [W, W], rot=angle)
            centers[v, b], scales[v, b], angles[v, b], flips[v, b] = center, scale.item(), angle.item(), flip
    save("viewkps", kps=kps, centers=centers, scales=scales, angles=angles, flips=flips, mats=mats, out=out, img_w=np.int64(W))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mixunc":
        mixunc_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "viewkps":
        viewkps_fixture()
        sys.exit(0)
    business_fixture()
    mixunc_fixture()
    viewkps_fixture()
    chain_fixture("chain_mt", B=2, K=3, J=3, H=64, W=64, M=1, seed=1388)
    chain_fixture("chain_dual", B=4, K=4, J=5, H=32, W=32, M=2, seed=1389)
    decode_fixture()
    render_fixture()
    loss_fixture()
    ema_fixture()
