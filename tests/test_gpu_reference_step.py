"""One synthetic training step of the reference's drivers, run twice on CUDA tensors: with the reference's OWN
criteria / EMA (imported unmodified through oracle/ref_import.py from /root/reference or, on the GPU box, from the
staged copy under baseline/_ref, tools/stage_reference.py) and again after `ubpl_b200.install.install()` patched
the reference's modules.  The loops below are the drivers' own statements restated on synthetic model outputs:

  MT_UBPL          projects/MT_UBPL.py:246-298 (mtc / pec / epc criterion loops), :333-338
                   (total.backward(retain_graph=True) per model, update_ema_variables)
  DualPose_UBPL    projects/DualPose_UBPL.py:201-244 (JointDistLoss_mt2 consistency, JointPseudoLoss3 ensemble
                   pseudo), :281 (EMA)

Bars (north star): python-int counts exact, losses / gradients / joint scores within 1e-5 relative, EMA bit-exact.
Skipped where no reference tree is available."""
import types

import numpy as np
import pytest
import torch

import ref_import

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.reference_available(), reason="no reference tree (run tools/stage_reference.py)")]

RTOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return ref_import.load_reference()


def _blobs(shape, gen, dev, amp_lo=0.6, amp_hi=1.15):
    """Heat-map-like tensors [..., H, W]: one Gaussian blob per map plus noise; some maxima cross scoreThr = 0.95."""
    *lead, H, W = shape
    cx = torch.rand(*lead, 1, 1, generator=gen) * (W - 16) + 8
    cy = torch.rand(*lead, 1, 1, generator=gen) * (H - 16) + 8
    amp = amp_lo + (amp_hi - amp_lo) * torch.rand(*lead, 1, 1, generator=gen)
    ys = torch.arange(H).view(H, 1).float()
    xs = torch.arange(W).view(1, W).float()
    t = amp * torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 18.0) + 0.02 * torch.randn(*lead, H, W, generator=gen)
    return t.to(dev)


class _Tiny(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.a = torch.nn.Conv2d(3, 8, 3)
        self.bn = torch.nn.BatchNorm2d(8)
        self.b = torch.nn.Linear(33, 17)


def _inputs(M, K, B, S, J, H, W, dev, seed):
    g = torch.Generator().manual_seed(seed)
    # students and teachers of both models look at the same key points: one centre per (view, sample, joint), integer
    # (so that the blob's peak is its amplitude), amplitudes and noise per map
    cx = torch.randint(8, W - 8, (1, K, B, 1, J, 1, 1), generator=g).float()
    cy = torch.randint(8, H - 8, (1, K, B, 1, J, 1, 1), generator=g).float()
    ys = torch.arange(H).view(H, 1).float()
    xs = torch.arange(W).view(1, W).float()
    blob = torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 18.0)

    def maps():
        amp = 0.7 + 0.45 * torch.rand(M, K, B, S, J, 1, 1, generator=g)
        return (amp * blob + 0.01 * torch.randn(M, K, B, S, J, H, W, generator=g)).to(dev)
    d = dict(outs=maps(), outs_ema=maps(), heat=_blobs((K, B, J, H, W), g, dev, 1.0, 1.0))
    d["gate"] = (torch.rand(K, B, J, generator=g) < 0.7).float().to(dev)
    islabeled = torch.zeros(B, dtype=torch.bool)
    islabeled[B - B // 2:] = True                                          # TwoStream layout (utils/mt/data.py:121-129)
    d["islabeled"] = islabeled.to(dev)
    return d


def _weights(ref, islabeled, dev):
    """projects/tools.py:14-31 through the reference's own helper (it is host glue that stays with the reference)."""
    args = types.SimpleNamespace(device=dev, pseudoWeight=1.0)
    return ref.tools.ProjectTools.getSampleWeight([islabeled], args), ref.tools.ProjectTools.getSampleWeight_nega([islabeled], args)


def _mt_ubpl_step(L, upd, d, ref, dev, S, epo):
    """The statements of projects/MT_UBPL.py:246-298,333-338 on the tensors of `d` with the criteria of module L."""
    outs = d["outs"].clone().requires_grad_(True)
    outs_ema = d["outs_ema"].clone()
    M, K = outs.shape[:2]
    args = types.SimpleNamespace(consWeight=0.7, poseWeight=1.3, ensemblePseudoWeight=0.9, pseudoScoreThr=0.95, epo=epo,
                                 ema_decay=0.999)
    consistency = L.JointDistLoss(nStack=1)
    pose = L.JointMSELoss(nStack=S, useKPsGate=True, useSampleWeight=True)
    pseudo2 = L.JointPseudoLoss3(nStack=S, scoreThr=args.pseudoScoreThr)
    sw, nega = _weights(ref, d["islabeled"], dev)
    rec = {}
    mtc_losses, pec_losses, epc_losses = [], [], []
    for m in range(M):
        s_, c_ = 0., 0
        for a in range(K):
            loss, n = consistency(outs[m, a, :, -1], outs_ema[m, a, :, -1])
            s_ += loss
            c_ += n
        mtc_losses.append(args.consWeight * ((s_ / c_) if c_ > 0 else s_))
        rec["mtc_count_%d" % m] = c_
    for m in range(M):
        s_, c_ = 0., 0
        for a in range(K):
            loss, n = pose(outs[m, a], d["heat"][a], d["gate"][a], sw[0])
            s_ += loss
            c_ += n
        pec_losses.append(args.poseWeight * ((s_ / c_) if c_ > 0 else s_))
        rec["pec_count_%d" % m] = c_
    n_pseudo, n_sel, scores = 0, 0, []
    for m in range(M):
        s_, c_ = 0., 0
        for a in range(K):
            loss, n, ns, js, t1, t2 = pseudo2(outs[m, a], outs_ema.clone()[:, a].detach(), nega[0])
            s_ += loss
            c_ += n
            n_pseudo += n
            n_sel += ns
            scores.append(js)
        epc_losses.append(args.ensemblePseudoWeight * ((s_ / c_) if c_ > 0 else s_))
    rec["n_pseudo"], rec["n_sel"] = n_pseudo, n_sel
    rec["score"] = torch.stack(scores, 0).mean(0).detach().cpu().numpy()
    for m in range(M):
        total = pec_losses[m] + mtc_losses[m] + epc_losses[m]
        total.backward(retain_graph=True)
        rec["total_%d" % m] = float(total.detach())
    rec["grad"] = outs.grad.detach().cpu().numpy()
    models = [_Tiny(10 + m).to(dev) for m in range(M)]
    emas = [_Tiny(20 + m).to(dev) for m in range(M)]
    for m in range(M):
        upd(models[m], emas[m], args)
    rec["ema"] = [p.detach().cpu().numpy() for e in emas for p in e.parameters()]
    rec["ema_buffers"] = [b.detach().cpu().numpy() for e in emas for b in e.buffers()]
    return rec


def _dualpose_step(L, upd, d, ref, dev, S, epo):
    """projects/DualPose_UBPL.py:201-244,281: one student view, one weakly augmented teacher view per model."""
    outs = d["outs"][:, 0].clone().requires_grad_(True)                    # [M,B,S,J,H,W]
    outs_ema = d["outs_ema"][:, 0].clone()
    M = outs.shape[0]
    args = types.SimpleNamespace(consWeight=0.8, ensemblePseudoWeight=1.1, pseudoScoreThr=0.95, epo=epo, ema_decay=0.999,
                                 device=dev, pseudoWeight=1.0)
    cons = L.JointDistLoss_mt2(nStack=1, useKPsGate=False, useSampleWeight=True, scoreThr=args.pseudoScoreThr)
    pseudo2 = L.JointPseudoLoss3(nStack=S, scoreThr=args.pseudoScoreThr)
    w_cons = ref.tools.ProjectTools.getSampleWeight_mt_cons(d["islabeled"], args)
    w_nega = ref.tools.ProjectTools.getSampleWeight_mt_nega(d["islabeled"], args)
    rec = {}
    losses = []
    for m in range(M):
        loss, n, n_pseudo, n_sel, score = cons(outs[m, :, -1], outs_ema[m, :, -1], sampleWeight=w_cons)
        mtc = args.consWeight * ((loss / n) if n > 0 else loss)
        rec["cons_%d" % m] = (n, n_pseudo, n_sel)
        rec["cons_score_%d" % m] = score.detach().cpu().numpy()
        loss2, n2, ns2, js, t1, t2 = pseudo2(outs[m], outs_ema.clone().detach(), w_nega)
        epc = args.ensemblePseudoWeight * ((loss2 / n2) if n2 > 0 else loss2)
        rec["epc_%d" % m] = (n2, ns2)
        rec["epc_score_%d" % m] = js.detach().cpu().numpy()
        losses.append(mtc + epc)
    for m in range(M):
        losses[m].backward(retain_graph=True)
        rec["total_%d" % m] = float(losses[m].detach())
    rec["grad"] = outs.grad.detach().cpu().numpy()
    models = [_Tiny(30 + m).to(dev) for m in range(M)]
    emas = [_Tiny(40 + m).to(dev) for m in range(M)]
    for m in range(M):
        upd(models[m], emas[m], args)
    rec["ema"] = [p.detach().cpu().numpy() for e in emas for p in e.parameters()]
    rec["ema_buffers"] = [b.detach().cpu().numpy() for e in emas for b in e.buffers()]
    return rec


def _compare(a, b):
    assert set(a) == set(b)
    for k in a:
        if k in ("ema", "ema_buffers"):
            for x, y in zip(a[k], b[k]):
                assert np.array_equal(x, y), k                              # EMA bit-exact, BN buffers untouched
        elif isinstance(a[k], np.ndarray):
            np.testing.assert_allclose(b[k], a[k], rtol=RTOL, atol=1e-9, err_msg=k)
        elif isinstance(a[k], float):
            np.testing.assert_allclose(b[k], a[k], rtol=RTOL, err_msg=k)
        else:
            assert a[k] == b[k], (k, a[k], b[k])                            # python-int counts exact


@pytest.mark.parametrize("driver", ["MT_UBPL", "DualPose_UBPL"])
def test_driver_step_reference_vs_installed(ref, driver):
    import importlib
    import ubpl_b200  # noqa: F401
    from ubpl_b200 import install
    dev = torch.device("cuda", 0)
    S, J, H, W = 2, 6, 64, 64
    d = _inputs(M=2, K=2, B=6, S=S, J=J, H=H, W=W, dev=dev, seed=7 if driver == "MT_UBPL" else 8)
    step = _mt_ubpl_step if driver == "MT_UBPL" else _dualpose_step
    losses_mod = importlib.import_module("utils.losses")
    params_mod = importlib.import_module("utils.parameters")
    ref_cls = losses_mod.JointPseudoLoss3
    want = step(losses_mod, params_mod.update_ema_variables, d, ref, dev, S, epo=3)          # the reference's own classes
    try:
        done = install.install()
        assert "utils.losses.JointPseudoLoss3" in done
        assert losses_mod.JointPseudoLoss3 is not ref_cls                                  # really patched
        got = step(losses_mod, params_mod.update_ema_variables, d, ref, dev, S, epo=3)
    finally:
        install.uninstall()
    assert losses_mod.JointPseudoLoss3 is ref_cls
    _compare(want, got)
    assert (want["n_sel"] if driver == "MT_UBPL" else want["epc_0"][1]) > 0                 # some joints pass the score masks
