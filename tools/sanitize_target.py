"""Small workload for compute-sanitizer (tools/sanitize.sh): __graft_entry__.smoke() plus one fused step per path --
fixed and quantile selection, one and two teachers, maps that take the cooperative exhaustive decode (white noise,
NaN, constant, singular theta), the multi-rank selector with its ranks emulated on this GPU, the dense-target loss,
the feature covariance and the EMA.  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as G  # noqa: E402

G.smoke()
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import ops, pipeline, synth  # noqa: E402

for (M, sel) in ((1, "fixed"), (1, "quantile"), (2, "fixed"), (2, "quantile")):
    d = synth.make_batch(B=6, K=4, J=5, M=M, S=2, seed=11 + M, device="cuda", noise_only_frac=1.0, neg_frac=0.3)
    t = d["teacher"]
    t[0, 0, 0, 0] = torch.randn(64, 64, device="cuda") * 0.02          # structure-less: exhaustive
    t[0, 1, 1, 1, 20, 30] = float("nan")
    t[0, 2, 2, 2] = 0.25                                              # constant
    d["theta"][3, 3] = 0.0                                             # singular transform
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select=sel, distThrMax=2.0)
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = pipeline.pseudo_label_step(t, d["student"], d["theta"], d["flip"], dec, w, cfg, stats=stats)
    torch.cuda.synchronize()
    print("step M=%d %s: loss %.6g, exhaustive maps %d of %d" % (M, sel, float(r["summary"][0]) * float(r["grad_scale"]),
                                                                int(stats[0]), int(stats[2])), flush=True)
rng = np.random.default_rng(5)
dist = torch.from_numpy(np.round(rng.gamma(2.0, 2.0, (3, 500)) * 4) / 4).cuda()
leg = torch.from_numpy((rng.random((3, 500)) > 0.1).astype(np.uint8)).cuda()
r = ops.select_quantile_emul(dist, leg, 7, 0.0, 0.5, 1.0, n_per_rank=[500, 37, 499])
print("emulated 3-rank selector: status", r["status"].tolist(), "thr", r["thr"].tolist(), flush=True)
pred = torch.rand(4, 2, 5, 64, 64, device="cuda")
tgt = torch.rand(2, 4, 2, 5, 64, 64, device="cuda")
r = ops.dense_mse(pred, tgt[:, :, -1], mask_mode=1, thr=0.95)
f1 = torch.randn(2, 2, 8, 32, 32, device="cuda")
r = ops.features_cov(f1, f1.flip(0))
e = [torch.randn(1000, device="cuda"), torch.randn(7, 3, device="cuda")]
p = [torch.randn(1000, device="cuda"), torch.randn(7, 3, device="cuda")]
ops.EmaPlan(p, e).step(0.75)
torch.cuda.synchronize()
print("sanitize target done", flush=True)
