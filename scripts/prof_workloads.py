#!/usr/bin/env python
"""prof_workloads.py -- one small, named workload per kernel family, for `ncu -k regex:<kernel>` captures (scripts/gpu_profiles_r2.sh).

Every workload runs twice (first = warm-up: JIT-free, but it loads columns and sizes scratch), so captures use `-s` to skip the
first pass' launches where that matters."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def table(name, spec, rows, seed):
    from oracle import oracle as O
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(base, f"dfdb_b200_prof_{name}_{rows}")
    if not os.path.exists(os.path.join(path, ".complete")):
        shutil.rmtree(path, ignore_errors=True)
        O.gen_table(path, spec, rows, 65536, seed, os.cpu_count() or 1)
        open(os.path.join(path, ".complete"), "w").write("ok")
    return path


def main():
    what = sys.argv[1]
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    import numpy as np
    import torch
    import dfdb_b200 as D
    from dfdb_b200 import R, _capi
    torch.cuda.set_device(0)
    _capi.init(0)
    L = _capi.lib()
    if what in ("lane", "walker"):
        if what == "lane":
            L.dfdb_set_option(b"lz4_flavour", 3)
        t = D.open_table(table("ab", "a:Int64:iuniform:1:100;b:Float64:funiform", rows, 0xDFDB0002), mode=D.LOAD_HBM)
        v = t[(t.a > 25) & (t.a <= 75), ["b"]]
        for _ in range(2):
            D.aggregate(v.b)
    elif what == "strings":          # vm_mask_kernel on strings, gather_strings_kernel, str_offsets_kernel
        t = D.open_table(table("s", "s:String:brands;k:Int64:iseq", rows, 0xDFDB0003), mode=D.LOAD_HBM)
        for _ in range(2):
            D.materialize(t[t.s == "sony", ["s"]])
    elif what == "missings":         # fused_scan_kernel (mask out), gather_fixed_kernel on nullable columns
        t = D.open_table(table("m", "a:Missing(Int64):iuniform:1:100:m=0.1;b:Missing(Float64):funiform:m=0.1;d:Float64:funiform", rows, 0xDFDB0004), mode=D.LOAD_HBM)
        for _ in range(2):
            D.materialize(t[D.coalesce(t.a > 50, False) & D.coalesce(t.b < 0.5, False), ["a", "b", "d"]])
    elif what == "range":            # range_stage_kernel: a range stage behind a predicate ranks rows among all survivors
        t = D.open_table(table("ab", "a:Int64:iuniform:1:100;b:Float64:funiform", rows, 0xDFDB0002), mode=D.LOAD_HBM)
        for _ in range(2):
            D.nrow(t[t.a > 50, :][R(10, 3, rows // 4), ["b"]])
    elif what == "arith":            # vm_mask_kernel on arithmetic, agg_vm_kernel
        t = D.open_table(table("ab", "a:Int64:iuniform:1:100;b:Float64:funiform", rows, 0xDFDB0002), mode=D.LOAD_HBM)
        for _ in range(2):
            D.aggregate((t[(t.a * 3 + 1) % 7 == 0, :].b * 2.0))
    elif what == "group":            # group_reduce_kernel
        t = D.open_table(table("gs", "s:String:brands;a:Int64:iuniform:1:100;b:Float64:funiform", rows, 0xDFDB0005), mode=D.LOAD_HBM)
        for _ in range(2):
            D.groupreduce(t[t.a > 50, :], ["s"], total="b", n="a")
    elif what == "zone":             # zone_map_kernel + a pruned scan
        p = table("q", "q:Int64:iseq;b:Float64:funiform", rows, 0xDFDB0006)
        for f in os.listdir(p):
            if f.endswith(".zmap"):
                os.unlink(os.path.join(p, f))
        t = D.open_table(p, mode=D.LOAD_HBM)
        t.build_zonemaps()
        for _ in range(2):
            D.aggregate(t[(t.q > rows // 2) & (t.q <= rows // 2 + rows // 100), ["b"]].b)
    elif what == "write":            # pack_bodies_kernel, lz4_compress_kernel, compact_payloads_kernel
        rng = np.random.default_rng(1)
        n = min(rows, 50_000_000)
        data = {"a": rng.integers(1, 101, n).astype(np.int64), "f": rng.random(n)}
        base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        for k in range(2):
            p = os.path.join(base, f"dfdb_b200_prof_write_{k}")
            shutil.rmtree(p, ignore_errors=True)
            D.create_table(p, data).close()
            shutil.rmtree(p, ignore_errors=True)
    else:
        raise SystemExit(f"unknown workload {what}")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
