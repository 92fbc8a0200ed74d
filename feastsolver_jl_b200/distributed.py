"""torch.distributed plumbing: one process per GPU; the ncclUniqueId created by the
library on rank 0 is broadcast over the existing process group and every rank
attaches its context (feast_comm_init).  The data-path collective itself
(ncclAllReduce of the Q accumulator) runs inside libfeast_cuda.so on its stream."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .partition import node_owners


def make_comm_hook(contour_nodes=None, balanced=True):
    """Returns a callable(ctx) for the `comm=` argument of feast/gen_feast/nlfeast."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    world, rank = dist.get_world_size(), dist.get_rank()

    def hook(ctx):
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(ctx.device)   # the hook may run on a helper thread: the current device is per thread
        lib = _lib.load()
        payload = [None]
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(lib.feast_comm_unique_id(buf))
            payload = [bytes(buf.raw)]
        dist.broadcast_object_list(payload, src=0)
        ctx.comm_init(world, rank, payload[0])
        hook.ctx_ranks = (world, rank)

    def owners_hook(ctx, nodes):
        if balanced:
            ctx.set_node_owners(node_owners(nodes, world))

    hook.set_owners = owners_hook
    return hook
