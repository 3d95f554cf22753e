"""Import shim for the UNMODIFIED reference (test infrastructure, build container only).

`/root/reference` is a pure-Python repo whose hot-path modules import a few packages that are
not installed here (skimage, openpyxl, imageio, matplotlib) and one module evaluates
`torch.Tensor([...]).cuda()` in a default argument at import time
(/root/reference/utils/udaap/imutils.py:190).  None of the stubbed names is touched by the
hot-path functions the oracle restates.  This shim is used ONLY to
  * pin `oracle/ubpl_oracle.py` against the reference's own functions (tests/test_oracle_vs_reference.py),
  * generate the committed fixtures under tests/golden/ (tests/golden/make_golden.py).
On the GPU box (where /root/reference is not mounted) it resolves to the staged, unmodified copy under
baseline/_ref/ (tools/stage_reference.py; git-ignored) and additionally serves bench.py's CPU baseline
(oracle/ref_chain.py).  Nothing in the product package imports it.
"""
import os
import sys
import types
import contextlib

# /root/reference in the build container; on the GPU box the byte-for-byte staged copy that
# tools/stage_reference.py leaves under the git-ignored baseline/_ref/
_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("UBPL_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/utils") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


@contextlib.contextmanager
def _cuda_is_noop():
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


_cache = {}


def load_reference():
    """Returns a namespace with the reference's hot-path classes/functions (imported, not copied)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _stub("openpyxl")
    _stub("openpyxl.styles", PatternFill=object)
    sk = _stub("skimage")
    sk.transform = _stub("skimage.transform")
    sk.data = _stub("skimage.data")
    _stub("imageio")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with _cuda_is_noop():
        import importlib
        ns = types.SimpleNamespace()
        ns.losses = importlib.import_module("utils.losses")
        ns.augment = importlib.import_module("utils.augment")
        ns.process = importlib.import_module("utils.process")
        ns.evaluation = importlib.import_module("utils.evaluation")
        ns.business = importlib.import_module("utils.business")
        ns.parameters = importlib.import_module("utils.parameters")
        ns.udaap_eval = importlib.import_module("utils.udaap.evaluation")
        ns.udaap_tf = importlib.import_module("utils.udaap.transforms")
        ns.tools = importlib.import_module("projects.tools")
    ns.aug = ns.augment.AugmentUtils
    ns.proc = ns.process.ProcessUtils
    ns.eval = ns.evaluation.EvaluationUtils
    ns.bus = ns.business.BusinessUtils
    _cache["ns"] = ns
    return ns
