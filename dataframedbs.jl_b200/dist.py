"""Multi-GPU combine step (K8): one process per GPU, block-range shards, no data-path collective.

Each rank produces a `dfdb_agg` partial for its shard; the 88-byte structs are all-gathered with
torch.distributed (NCCL on GPUs, gloo in the CPU tests) and folded in rank order on every rank by
`dfdb_agg_fold`, so every rank holds the same, deterministically combined result.  Row counts of sharded
`nrow` / `materialize` calls combine the same way (sum / concatenation in rank order).
"""
from __future__ import annotations

import ctypes as C

from . import _capi

COMM_ID_BYTES = 128


def comm_init_from_torch(rank: int, world: int, device=None):
    """Create the LIBRARY's NCCL communicator (dfdb_comm_init): rank 0 makes the 128-byte id (dfdb_comm_unique_id) and
    torch.distributed only carries it to the other ranks.  After this the combine step runs inside the C ABI
    (dfdb_scan_aggregate_all: ncclAllGather on the scan stream + dfdb_agg_fold), with no torch call per scan."""
    import torch
    import torch.distributed as dist

    L = _capi.lib()
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    if rank == 0:
        _capi.check(L.dfdb_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(raw)
    _capi.check(L.dfdb_comm_init(rank, world, buf))


def comm_destroy():
    _capi.check(_capi.lib().dfdb_comm_destroy())


def allgather_fold(agg: _capi.Agg, device=None) -> _capi.Agg:
    """All-gather one dfdb_agg per rank and fold them in rank order.  `device`: torch device of the exchange buffers."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    n = C.sizeof(_capi.Agg)
    mine = torch.frombuffer(bytearray(bytes(agg)), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * n, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    raw = out.cpu().numpy().tobytes()
    parts = (_capi.Agg * world)(*[_capi.Agg.from_buffer_copy(raw[i * n:(i + 1) * n]) for i in range(world)])
    res = _capi.Agg()
    _capi.check(_capi.lib().dfdb_agg_fold(parts, world, C.byref(res)))
    return res


def allreduce_count(n: int, device=None) -> int:
    import torch
    import torch.distributed as dist

    t = torch.tensor([n], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def allgather_counts(n: int, device=None):
    """Every rank's integer, in rank order (the survivor counts of a sharded selection stage)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return [int(x) for x in out.cpu().tolist()]


def resolve_selection(view, device=None):
    """Multi-process flavour of api.resolve_sharded_selection: the survivor counts travel by all_gather."""
    from . import api

    api.resolve_sharded_selection([view], lambda n: allgather_counts(n, device))
