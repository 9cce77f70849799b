#!/usr/bin/env python
"""h2d_ceiling.py -- what the box gives: pinned host -> device bandwidth with 1..N ranks copying at the same time.

The transfer-inclusive mode of bench.py (`e2e`) is bound by this path; its 8-GPU efficiency in round 1 was 0.42 with nobody
having measured the node's aggregate ceiling.  Launch with torchrun (one rank per GPU).  Every rank copies its own 2 GiB of
page-locked memory (allocated after dfdb_init, i.e. under the NUMA policy the library sets) to its
GPU, all ranks at once; rank 0 prints one JSON line: per-rank GB/s, the aggregate, and each rank's NUMA node.
DFDB_NO_NUMA=1 gives the figure without the NUMA binding."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from dfdb_b200 import _capi
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _capi.init(local)
    L = _capi.lib()
    nbytes = 2 << 30
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)      # (after dfdb_init: allocated under the process's NUMA policy)
    host.fill_(1)                                                       # touch: pages are placed now
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    out = {}
    for label, active in (("all_ranks", world), ("one_rank", 1)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        if rank < active:
            e0.record(stream)
            for _ in range(reps):
                dev.copy_(host, non_blocking=True)
            e1.record(stream)
        torch.cuda.synchronize()
        gbs = (reps * nbytes / 1e9) / (e0.elapsed_time(e1) / 1e3) if rank < active else 0.0
        t = torch.tensor([gbs, float(L.dfdb_numa_node())], dtype=torch.float64, device="cuda")
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
        else:
            g = [t]
        out[label] = {"per_rank_gbs": [round(float(x[0]), 2) for x in g], "aggregate_gbs": round(sum(float(x[0]) for x in g), 2),
                      "numa_node": [int(x[1]) for x in g]}
    if rank == 0:
        print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "numa_binding": not os.environ.get("DFDB_NO_NUMA"), **out}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
