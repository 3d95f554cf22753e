"""Runs a few pseudo-label steps of a bench config (default c2) for ncu: `python tools/prof_step.py [cfg] [steps]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import ubpl_b200  # noqa: E402,F401
from ubpl_b200 import ops, pipeline, synth  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
c = bench.CONFIGS[cfgname]
d = synth.make_batch(B=c["B"], K=c["K"], J=c["J"], H=c["H"], W=c["W"], M=c["M"], S=c["S"], seed=1388, device="cuda")
dec = ops.decode_coeffs(d["center"], d["scale"], [c["H"], c["W"]])
w = pipeline.nega_weights(d["islabeled"], 1.0)
cfg = pipeline.StepConfig(select=c["select"], distThrMax=bench.DIST_THR_MAX)
shapes = json.load(open(os.path.join(ROOT, "ubpl-poseestimation_b200", "hg_param_shapes.json")))[c["hg"]]
params = [torch.randn(*s, device="cuda") * 0.02 for s in shapes]
emas = [torch.randn(*s, device="cuda") * 0.02 for s in shapes]
plan = ops.EmaPlan(params, emas)
stats = torch.zeros(4, dtype=torch.int64, device="cuda")
for _ in range(steps):
    r = pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg, stats=stats, ema=plan, alpha=0.75)
torch.cuda.synchronize()
s = stats.tolist()
print("maps", s[2], "exhaustive", s[0], "evaluated px per map", s[1] / max(1, s[2]), "loss", float(r["summary"][0]) * float(r["grad_scale"]))
