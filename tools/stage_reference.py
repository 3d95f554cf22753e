"""Stages the UNMODIFIED reference's hot-path sources for the GPU box.

`/root/reference` is mounted in the build container only; `gpurun` ships `/root/repo`.  This script copies
the reference's `utils/`, `models/`, `projects/` and `GLOB.py` (pure Python, ~400 KB, no data) byte for byte
into `baseline/_ref/`, which is git-ignored (nothing of the reference enters the history) but not
gpurun-ignored (so it travels).  Users of the staged tree: `oracle/ref_import.py` (falls back to it when
`/root/reference` is absent), hence `bench.py --impl reference` / `cpu_baseline` (kind "reference": the
reference's own functions timed on the box's host cores), `tests/test_oracle_vs_reference.py` and
`tests/test_gpu_reference_step.py` (the reference's criterion loops on CUDA tensors, with and without
`ubpl_b200.install()`).  The product package never imports it.

    python tools/stage_reference.py            # copy (idempotent); prints what it did
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("UBPL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["utils", "models", "projects", "GLOB.py"]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "utils")):
        if verbose:
            print("stage_reference: %s is not mounted; nothing staged" % SRC)
        return False
    os.makedirs(DST, exist_ok=True)
    n = 0
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if os.path.isdir(s):
            for dirpath, dirnames, filenames in os.walk(s):
                dirnames[:] = [x for x in dirnames if x != "__pycache__"]
                rel = os.path.relpath(dirpath, s)
                os.makedirs(os.path.join(d, rel), exist_ok=True)
                for f in filenames:
                    if not f.endswith(".py"):
                        continue
                    sf, df = os.path.join(dirpath, f), os.path.join(d, rel, f)
                    if not (os.path.exists(df) and filecmp.cmp(sf, df, shallow=False)):
                        shutil.copyfile(sf, df)
                        n += 1
        elif os.path.isfile(s):
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print("stage_reference: %d file(s) copied into %s" % (n, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
