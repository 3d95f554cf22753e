"""Writes ubpl-poseestimation_b200/hg_param_shapes.json: the parameter shapes (in `.parameters()`
order) of the reference's StackedHourglass for the benchmark's EMA leg, obtained by instantiating
the UNMODIFIED reference model class (models/pose/hourglass.py:60-90) in the build container.
Shapes only -- no weights, no code."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402

ref_import.load_reference()
from models.pose.hourglass import StackedHourglass  # noqa: E402

out = {}
for name, (k, nstack) in {"hg2_j14": (14, 2), "hg2_j9": (9, 2), "hg2_j17": (17, 2), "hg2_j32": (32, 2)}.items():
    m = StackedHourglass(k, nstack, "AvgPool") if True else None
    shapes = [list(p.shape) for p in m.parameters()]
    out[name] = shapes
    print(name, len(shapes), sum(int(__import__("numpy").prod(s)) for s in shapes))
json.dump(out, open(os.path.join(ROOT, "ubpl-poseestimation_b200", "hg_param_shapes.json"), "w"))
