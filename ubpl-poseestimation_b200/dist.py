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
