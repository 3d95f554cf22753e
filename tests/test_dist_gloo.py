"""N>1 control flow of the global-quantile selector (ops.select_quantile with a process group) on
CPU: world_size-2 gloo, each rank owns half of the items, the extrema and the four key histograms
are all-reduced, and both ranks must arrive at the threshold / masks the oracle computes on the
concatenated data.  The five device primitives are replaced by a numpy stand-in (test
infrastructure; the product path always runs the CUDA kernels, covered by tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import ubpl_oracle as O


class NumpySelectBackend:
    """numpy statement of the K2 selection primitives (csrc/uncertainty_select.cu)."""

    def prepare(self, dist, legal):
        return dist.reshape(-1).double().contiguous(), legal.reshape(-1).double().contiguous()

    def extrema(self, dist):
        d = dist.numpy()
        ok = d < 999
        mx = d[ok & (d > 0)].max() if (ok & (d > 0)).any() else 0.0
        mn = min(999.0, d[ok].min()) if ok.any() else 999.0
        return torch.tensor([mx, mn], dtype=torch.float64)

    def reliability(self, dist, legal, ext, reliableDistMin):
        dmax, dmin = float(ext[0]), float(ext[1])
        if dmax == 0:
            dmax = 999.0
        if dmin > reliableDistMin:
            dmin = reliableDistMin
        d = dist.numpy()
        e2 = np.where(d != 999, d, dmax)
        unc = np.where(legal.numpy() > 0, (e2 - dmin) / (dmax - dmin), 1.0)
        rel = 1.0 - unc
        b = rel.view(np.uint64)
        keys = np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))
        return torch.from_numpy(rel), torch.from_numpy(keys.view(np.int64))

    def state(self, k, device):
        return torch.zeros(1, dtype=torch.int64), torch.full((1,), k, dtype=torch.int64), torch.zeros(65536, dtype=torch.int32)

    def histogram(self, keys, prefix, shift, hist):
        k = keys.numpy().view(np.uint64)
        if shift < 48:
            pf = np.uint64(prefix.numpy().view(np.uint64)[0])
            k = k[(k >> np.uint64(shift + 16)) == (pf >> np.uint64(shift + 16))]
        bins = ((k >> np.uint64(shift)) & np.uint64(0xffff)).astype(np.int64)
        hist.copy_(torch.from_numpy(np.bincount(bins, minlength=65536).astype(np.int32)))

    def descend(self, hist, shift, prefix, k_rem):
        h = hist.numpy().astype(np.int64)[::-1]
        c = np.cumsum(h)
        k = int(k_rem[0])
        pos = int(np.searchsorted(c, k, side="right"))
        pos = min(pos, 65535)
        b = 65535 - pos
        before = int(c[pos - 1]) if pos > 0 else 0
        pf = int(prefix.numpy().view(np.uint64)[0]) if shift < 48 else 0
        pf = (pf & ~(0xffff << shift)) | (b << shift)
        prefix.copy_(torch.from_numpy(np.array([pf], np.uint64).view(np.int64)))
        k_rem[0] = k - before

    def apply(self, rel, J, prefix, reliableThr):
        k = np.uint64(prefix.numpy().view(np.uint64)[0])
        bits = (k & np.uint64((1 << 63) - 1)) if (k >> np.uint64(63)) else ~k
        kth = float(np.array([bits], np.uint64).view(np.float64)[0])
        thr = max(reliableThr, kth)
        en = rel.numpy() > thr
        counts = np.zeros(J + 1, np.int32)
        for i in np.nonzero(en)[0]:
            counts[i % J] += 1
            counts[J] += 1
        return (torch.from_numpy(en.astype(np.uint8)), torch.from_numpy(en.astype(np.float32)), torch.from_numpy(counts),
                torch.tensor([thr], dtype=torch.float64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dist_all, legal_all, J, pct, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    import ubpl_b200  # noqa: F401
    from ubpl_b200 import ops
    n = dist_all.numel() // world
    sl = slice(rank * n, (rank + 1) * n)
    r = ops.select_quantile(dist_all[sl], legal_all[sl], J, 0.0, pct, 1.0, group=td.group.WORLD, backend=NumpySelectBackend())
    out[rank] = (float(r["thr"]), r["enable"].numpy().copy(), r["reliability"].numpy().copy())
    td.destroy_process_group()


@pytest.mark.parametrize("pct", [0.5, 0.1])
def test_global_quantile_two_ranks_gloo(pct):
    rng = np.random.default_rng(17)
    J, n = 6, 6 * 40
    d = np.round(rng.gamma(2.0, 3.0, n) * 4) / 4
    d[rng.random(n) < 0.2] = 999.0
    d[:n // 2] *= 0.5                              # the two shards have different ranges
    d[d > 900] = 999.0
    legal = (rng.random(n) < 0.9).astype(np.float64)
    rel, thr, en = O.filter_dual(d, legal, 0.0, pct, 1.0)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, torch.from_numpy(d), torch.from_numpy(legal), J, pct, out), nprocs=2, join=True)
    for rank in (0, 1):
        t, e, r = out[rank]
        sl = slice(rank * n // 2, (rank + 1) * n // 2)
        assert t == thr                                              # the same global threshold on both ranks
        assert np.array_equal(e.astype(bool), en[sl])                 # masks bit-exact
        assert np.array_equal(r, rel[sl])


def test_single_rank_numpy_backend_matches_oracle():
    import ubpl_b200  # noqa: F401
    from ubpl_b200 import ops
    rng = np.random.default_rng(3)
    d = np.round(rng.gamma(2.0, 3.0, 500) * 4) / 4
    legal = (rng.random(500) < 0.9).astype(np.float64)
    for pct in (0.0, 0.3, 0.5, 0.99, 1.0):
        rel, thr, en = O.filter_dual(d, legal, 0.0, pct, 1.0)
        r = ops.select_quantile(torch.from_numpy(d), torch.from_numpy(legal), 5, 0.0, pct, 1.0, backend=NumpySelectBackend())
        assert float(r["thr"]) == thr and np.array_equal(r["enable"].numpy().astype(bool), en)
