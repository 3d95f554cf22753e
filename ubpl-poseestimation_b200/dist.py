"""Multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous): creates the library's
own NCCL communicator used by the fused global-quantile selector (ubpl_select_quantile_dist)."""
import ctypes

import torch

from . import _lib


def init_nccl(group=None):
    """Collective: every rank of `group` (default WORLD) must call it.  Rank 0 creates an NCCL unique id,
    it is broadcast through torch.distributed, and each rank joins the communicator."""
    import torch.distributed as td
    group = group if group is not None else td.group.WORLD
    world, rank = td.get_world_size(group), td.get_rank(group)
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        _lib.call("ubpl_nccl_unique_id", ctypes.cast(buf, ctypes.c_void_p))
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    td.broadcast(t, src=td.get_global_rank(group, 0), group=group)
    raw = bytes(t.cpu().tolist())
    idbuf = (ctypes.c_char * 128).from_buffer_copy(raw)
    _lib.call("ubpl_nccl_init", ctypes.cast(idbuf, ctypes.c_void_p), world, rank)
    return world


def nccl_ranks():
    return int(_lib.lib().ubpl_nccl_ranks())


def destroy_nccl():
    _lib.call("ubpl_nccl_destroy")


# ---- peer-memory exchange buffer of the fused selector (ubpl_select_quantile_fused, use_p2p = 1) ------------
_p2p_world = 0


def init_p2p(group=None, max_items=1 << 16):
    """Collective: every rank of `group` allocates its exchange buffer, the CUDA-IPC handles are all-gathered
    through torch.distributed and each rank maps its peers' buffers (NVLink peer access).  Returns True when
    the fused peer-memory selector is usable; False (after releasing everything) when IPC mapping is not
    possible on this system -- the NCCL selector then remains in use."""
    global _p2p_world
    import torch.distributed as td
    group = group if group is not None else td.group.WORLD
    world, rank = td.get_world_size(group), td.get_rank(group)
    if world < 2 or world > 16:
        return False
    ok = 1
    buf = (ctypes.c_char * 64)()
    try:
        _lib.call("ubpl_p2p_alloc", world, int(max_items), ctypes.cast(buf, ctypes.c_void_p))
    except _lib.UbplError:
        ok = 0
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(bytes(buf)) + [ok], dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    td.all_gather(allh, mine, group=group)
    allh = [h.cpu() for h in allh]
    good = all(int(h[64]) == 1 for h in allh)
    if good:
        raw = b"".join(bytes(h[:64].tolist()) for h in allh)
        hb = (ctypes.c_char * (64 * world)).from_buffer_copy(raw)
        try:
            _lib.call("ubpl_p2p_open", ctypes.cast(hb, ctypes.c_void_p), world, rank)
        except _lib.UbplError:
            good = False
    flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN, group=group)
    if int(flag.item()) != 1:
        _lib.call("ubpl_p2p_close")
        _p2p_world = 0
        return False
    td.barrier(group=group)
    _p2p_world = world
    return True


def p2p_ready(group=None):
    """True when the exchange buffer of the fused selector is mapped for exactly the ranks of `group`."""
    if _p2p_world == 0:
        return False
    import torch.distributed as td
    group = group if group is not None else td.group.WORLD
    return td.get_world_size(group) == _p2p_world and int(_lib.lib().ubpl_p2p_ranks()) == _p2p_world


def p2p_status():
    return int(_lib.lib().ubpl_p2p_status())


def check_p2p():
    """Raises if a launch of the peer-memory selector gave up waiting for a peer (status word of the exchange
    buffer; one D2H copy).  The outputs of such a launch are NaN-poisoned (every mask 0) on the ranks that timed out
    while the late rank computes a valid result -- the ranks have diverged, so this must stop the run.  Call it where
    a sync is acceptable: every N steps, at epoch boundaries, before a checkpoint."""
    st = p2p_status()
    if st != 0:
        raise _lib.UbplError("ubpl_b200: the multi-GPU selector timed out waiting for a peer (status %d): the pseudo-label "
                             "masks of that step are void on this rank; raise UBPL_P2P_TIMEOUT_MS or use the NCCL "
                             "selector (UBPL_BENCH_P2P=0 / dist.init_nccl)" % st)


def destroy_p2p():
    global _p2p_world
    _lib.call("ubpl_p2p_close")
    _p2p_world = 0
