"""Multi-GPU parity of the peer-memory selector, run when the box has at least two GPUs (skipped otherwise; the
exchange protocol itself is also covered on one GPU by test_gpu_shapes.py::test_multirank_selector_emulated and the
N > 1 control flow on CPU by test_dist_gloo.py).  Spawns one rank per GPU with torchrun on tools/p2p_check.py, which
asserts on every rank that the selection with the digit histograms all-reduced over NVLink peer memory
(ubpl_select_quantile_fused, use_p2p = 1) == the single-GPU selection of the concatenated shards == the NCCL
histogram selector, bit for bit (thresholds, masks, reliabilities, gates), for five shard sizes x three quantiles."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 8])
def test_p2p_selector_across_gpus(nranks):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    port = 29500 + (os.getpid() % 400) + nranks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    out = r.stdout + r.stderr
    if "init_p2p: False" in out:
        pytest.skip("CUDA IPC peer mapping is not available on this box")
    assert r.returncode == 0, out[-3000:]
    assert "p2p selector mismatches over all ranks: 0" in out, out[-3000:]
