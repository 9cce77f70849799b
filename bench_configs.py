#!/usr/bin/env python
"""bench_configs.py -- the other single-GPU configurations of BASELINE.json (configs[2], configs[3]) on one B200.

Not the driver's bench line (that is bench.py = configs[1]); this script gives the same kind of numbers for the
string-filter and the missing-bearing multi-column configurations so that every kernel on the path has a measured
throughput beside its algorithmic bytes:

  config 3   200M rows, String column `s` (8 brands): `s .== "sony"` and `startswith.(s, "s")`, materialize the strings
  config 4   500M rows, a::Union{Int64,Missing}, b::Union{Float64,Missing}, c::Union{Int64,Missing}, d::Float64, 10 % missing:
             `coalesce.(a .> 50, false) .& coalesce.(b .< 0.5, false)`, materialize [a, b, c, d]

Per query: rows/s through the public API (open table resident in HBM -> decode -> predicate -> gather -> host arrays),
the device time of every phase (CUDA events inside the library), and a check of the result against the CPU oracle on a
prefix of the table.  One JSON line per query.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLOCK = 65536
CONFIGS = {
    3: dict(rows=200_000_000, seed=0xDFDB0003, spec="s:String:brands;k:Int64:iseq"),
    4: dict(rows=500_000_000, seed=0xDFDB0004,
            spec="a:Missing(Int64):iuniform:1:100:m=0.1;b:Missing(Float64):funiform:m=0.1;c:Missing(Int64):iuniform:1:100:m=0.1;d:Float64:funiform"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_table(cfg, rows, threads):
    from oracle import oracle as O
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(base, f"dfdb_b200_cfg_{rows}_{cfg['seed']:x}")
    marker = os.path.join(path, ".complete")
    if os.path.exists(marker):
        return path, json.load(open(marker))
    shutil.rmtree(path, ignore_errors=True)
    t0 = time.time()
    unc, comp = O.gen_table(path, cfg["spec"], rows, BLOCK, cfg["seed"], threads)
    info = {"rows": rows, "uncompressed": unc, "compressed": comp, "gen_s": round(time.time() - t0, 2)}
    json.dump(info, open(marker, "w"))
    log(f"[cfg] generated {rows} rows in {info['gen_s']} s: {unc / 1e9:.2f} GB -> {comp / 1e9:.2f} GB")
    return path, info


def phases(L):
    out = {}
    for name in ("decode", "unpack", "select", "consume", "d2h"):
        ms, n, b = C.c_double(), C.c_int64(), C.c_int64()
        L.dfdb_profile_get(name.encode(), C.byref(ms), C.byref(n), C.byref(b))
        out[name] = {"ms": round(ms.value, 3), "launches": n.value, "bytes": b.value}
    return out


def run_query(D, L, name, make_view, rows, reps, check):
    import torch
    v = make_view()
    fr = D.materialize(v)                       # warm-up (also loads the columns)
    times = []
    L.dfdb_profile_reset()
    L.dfdb_profile_enable(1)
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fr = D.materialize(v)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    L.dfdb_profile_enable(0)
    ph = phases(L)
    for k in ph:
        ph[k]["ms"] = round(ph[k]["ms"] / reps, 3)
        ph[k]["launches"] //= reps
        ph[k]["bytes"] //= reps
    best = min(times)
    nsel = fr.nrow()
    out_bytes = 0
    for n in fr.names:
        col = fr[n]
        if isinstance(col, D.FlatStringsVector):
            out_bytes += col.sizes.nbytes + len(col.data)
        else:
            out_bytes += getattr(col, "nbytes", 0) + (nsel if hasattr(col, "mask") else 0)
    line = {"query": name, "rows": rows, "selected": nsel, "ms": round(best * 1e3, 2), "rows_per_s": rows / best,
            "out_bytes": out_bytes, "phases_ms": {k: ph[k]["ms"] for k in ph},
            "device_ms": round(sum(ph[k]["ms"] for k in ph), 2),
            "decode_gbs": round(ph["decode"]["bytes"] / max(ph["decode"]["ms"], 1e-9) / 1e6, 1),
            "verified": check(v, fr)}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3, choices=[3, 4])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check-rows", type=int, default=2_000_000)
    args = ap.parse_args()
    import numpy as np
    import torch
    import dfdb_b200 as D
    from dfdb_b200 import R, _capi
    from oracle import oracle as O

    cfg = CONFIGS[args.config]
    rows = args.rows or cfg["rows"]
    threads = os.cpu_count() or 1
    path, info = ensure_table(cfg, rows, threads)
    torch.cuda.set_device(0)
    _capi.init(0)
    L = _capi.lib()
    t = D.open_table(path, mode=D.LOAD_HBM, device=0)
    ot = O.OracleTable(path)
    nchk = min(rows, args.check_rows)

    def check_prefix(make_prefixed):
        """Same query behind a leading range stage 1:nchk, GPU against the CPU oracle (bit-exact)."""
        def chk(_v, _fr):
            pv = make_prefixed()
            got = D.materialize(pv)
            exp = ot.materialize(D.plan_bytes(pv))
            ok = True
            for g, e in zip([got[n] for n in got.names], exp):
                if isinstance(g, D.FlatStringsVector):
                    ok &= bool(np.array_equal(g.sizes, e.sizes) and g.data == e.chars)
                elif isinstance(g, np.ma.MaskedArray):
                    ev, em = e
                    ok &= bool(np.array_equal(np.ma.getmaskarray(g), em) and np.array_equal(g.data[~em], ev[~em]))
                else:
                    ok &= bool(np.array_equal(g, e))
            return {"ok": ok, "prefix_rows": nchk, "prefix_selected": got.nrow()}
        return chk

    log(f"[cfg] config {args.config}: {rows} rows, {info['uncompressed'] / 1e9:.2f} GB decoded, {info['compressed'] / 1e9:.2f} GB compressed")
    if args.config == 3:
        run_query(D, L, 's .== "sony" -> materialize [s]', lambda: t[t.s == "sony", ["s"]], rows, args.reps,
                  check_prefix(lambda: t[R(1, nchk), :][t.s == "sony", ["s"]]))
        run_query(D, L, 'startswith.(s, "s") -> materialize [s, k]', lambda: t[D.startswith(t.s, "s"), ["s", "k"]], rows, args.reps,
                  check_prefix(lambda: t[R(1, nchk), :][D.startswith(t.s, "s"), ["s", "k"]]))
    else:
        pred = lambda: D.coalesce(t.ma > 50, False) & D.coalesce(t.mb < 0.5, False)   # noqa: E731
        names = [m.name for m in t.meta]
        a, b = names[0], names[1]
        pred = lambda: D.coalesce(getattr(t, a) > 50, False) & D.coalesce(getattr(t, b) < 0.5, False)   # noqa: E731
        run_query(D, L, "coalesce.(a .> 50, false) .& coalesce.(b .< 0.5, false) -> materialize [a, b, c, d]",
                  lambda: t[pred(), names], rows, args.reps, check_prefix(lambda: t[R(1, nchk), :][pred(), names]))
    t.close()
    ot.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
