#!/usr/bin/env python
"""decode_kinds.py -- K1 (LZ4 block decode) throughput per body kind on one B200.

One table, one column per body kind of src/io/blocks.jl (bits, Union{T,Missing}, String); every column is decoded on its
own by a count query that touches only it, and the library's own phase timer (CUDA events around the decode launch)
gives the kernel time.  Prints one JSON line per column: compressed / decoded bytes, ms, decoded GB/s and
(compressed + decoded) GB/s = the algorithmic bytes of K1.
Every column is also checked against the CPU oracle on a prefix (count of the same predicate).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SPEC = ("ia:Int64:iuniform:1:100;sq:Int64:iseq;fg:Float64:fgrid:1:0.1:2000;"
        "ma:Missing(Int64):iuniform:1:100:m=0.1;mb:Missing(Float64):funiform:m=0.1;"
        "s:String:brands;sd:String:decimal;ms:Missing(String):brands:m=0.1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cols", default="", help="columns to measure AND to generate (default: all eight)")
    ap.add_argument("--check-rows", type=int, default=1_000_000)
    args = ap.parse_args()
    import torch
    import dfdb_b200 as D
    from dfdb_b200 import R, _capi
    from oracle import oracle as O

    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    want_cols = [c for c in args.cols.split(",") if c]
    spec = ";".join(p for p in SPEC.split(";") if not want_cols or p.split(":")[0] in want_cols)
    path = os.path.join(base, f"dfdb_b200_kinds_{args.rows}" + ("_" + "_".join(want_cols) if want_cols else ""))
    if not os.path.exists(os.path.join(path, ".complete")):
        shutil.rmtree(path, ignore_errors=True)
        t0 = time.time()
        unc, comp = O.gen_table(path, spec, args.rows, 65536, 0xDFDB0007, os.cpu_count() or 1)
        open(os.path.join(path, ".complete"), "w").write("ok")
        print(f"[kinds] generated {args.rows} rows in {time.time() - t0:.1f} s: {unc / 1e9:.2f} GB -> {comp / 1e9:.2f} GB", file=sys.stderr, flush=True)
    torch.cuda.set_device(0)
    _capi.init(0)
    L = _capi.lib()
    t = D.open_table(path, mode=D.LOAD_HBM, device=0)
    ot = O.OracleTable(path)
    preds = {
        "ia": lambda tt: tt.ia > 50, "sq": lambda tt: tt.sq > 50, "fg": lambda tt: tt.fg > 1000.0,
        "ma": lambda tt: D.coalesce(tt.ma > 50, False), "mb": lambda tt: D.coalesce(tt.mb < 0.5, False),
        "s": lambda tt: tt.s == "sony", "sd": lambda tt: D.startswith(tt.sd, "-1"), "ms": lambda tt: D.coalesce(tt.ms == "sony", False),
    }
    want = [c for c in args.cols.split(",") if c] or list(preds)
    for name in want:
        meta = t.getmeta(name)
        comp, unc = C.c_int64(), C.c_int64()
        v = t[preds[name](t), [name]]
        n = D.nrow(v)                                            # warm-up, loads the column
        L.dfdb_table_column_stats(t._h, meta.id, C.byref(comp), C.byref(unc))
        L.dfdb_profile_reset()
        L.dfdb_profile_enable(1)
        for _ in range(args.reps):
            n = D.nrow(v)
        torch.cuda.synchronize()
        L.dfdb_profile_enable(0)
        ms, k, b = C.c_double(), C.c_int64(), C.c_int64()
        L.dfdb_profile_get(b"decode", C.byref(ms), C.byref(k), C.byref(b))
        dms = ms.value / args.reps
        nchk = min(args.rows, args.check_rows)
        pv = t[R(1, nchk), :][preds[name](t), [name]]
        ok = D.nrow(pv) == ot.count(D.plan_bytes(pv))
        print(json.dumps({"column": name, "type": meta.typestring, "rows": args.rows, "selected": n, "compressed": comp.value,
                          "decoded": unc.value, "ratio": round(unc.value / max(comp.value, 1), 2), "decode_ms": round(dms, 3),
                          "decoded_gbs": round(unc.value / max(dms, 1e-9) / 1e6, 1),
                          "alg_gbs": round((unc.value + comp.value) / max(dms, 1e-9) / 1e6, 1), "prefix_ok": bool(ok)}), flush=True)
        t.drop_decoded()
    t.close()
    ot.close()


if __name__ == "__main__":
    sys.exit(main())
