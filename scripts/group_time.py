#!/usr/bin/env python
"""group_time.py -- f4 (group-by reduce) wall time on one B200: few groups (8 brand strings) and many groups (Int64 1..1e6)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from prof_workloads import table


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
    import torch
    import dfdb_b200 as D
    from dfdb_b200 import _capi
    torch.cuda.set_device(0)
    _capi.init(0)
    t = D.open_table(table("gs", "s:String:brands;a:Int64:iuniform:1:100;b:Float64:funiform", rows, 0xDFDB0005), mode=D.LOAD_DECODED)
    t2 = D.open_table(table("gk", "k:Int64:iuniform:1:1000000;b:Float64:funiform", rows // 4, 0xDFDB0008), mode=D.LOAD_DECODED)
    for name, fn in (("8 groups (String key), a > 50, 2 value columns", lambda: D.groupreduce(t[t.a > 50, :], ["s"], total="b", n="a")),
                     ("100 groups (Int64 key), all rows, 1 value column", lambda: D.groupreduce(t[:, :], ["a"], total="b")),
                     ("1e6 groups (Int64 key), all rows, 1 value column", lambda: D.groupreduce(t2[:, :], ["k"], total="b"))):
        fn()
        torch.cuda.synchronize()
        L = _capi.lib()
        L.dfdb_profile_reset(); L.dfdb_profile_enable(1)
        t0 = time.time()
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        L.dfdb_profile_enable(0)
        import ctypes as C
        ph = {}
        for nm in ("decode", "select", "consume", "d2h"):
            ms, n, b = C.c_double(), C.c_int64(), C.c_int64()
            L.dfdb_profile_get(nm.encode(), C.byref(ms), C.byref(n), C.byref(b))
            ph[nm] = round(ms.value / 3, 2)
        print("   device phases (ms per call):", ph)
        ng = len(next(iter(r.values()))) if isinstance(r, dict) else -1
        print(f"{name}: {dt * 1e3:.1f} ms per call (wall, decoded columns cached), groups {ng}", flush=True)


if __name__ == "__main__":
    main()
