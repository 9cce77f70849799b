#!/usr/bin/env python
"""bench_configs.py -- the other single-GPU configurations of BASELINE.json (configs[2], configs[3]) on one B200.

Not the driver's bench line (that is bench.py = configs[1]); this script gives the same kind of numbers for the
string-filter and the missing-bearing multi-column configurations so that every kernel on the path has a measured
throughput beside its algorithmic bytes:

  config 3   200M rows, String column `s` (8 brands): `s .== "sony"` and `startswith.(s, "s")`, materialize the strings
  config 4   500M rows, a::Union{Int64,Missing}, b::Union{Float64,Missing}, c::Union{Int64,Missing}, d::Float64, 10 % missing:
             `coalesce.(a .> 50, false) .& coalesce.(b .< 0.5, false)`, materialize [a, b, c, d]

Per query: rows/s through the public API (open table resident in HBM -> decode -> predicate -> gather -> host arrays),
the device time of every phase (CUDA events inside the library), and a check of the WHOLE result against the CPU oracle:
the oracle materializes the same plan over block ranges on every host thread and reduces each output column to a 64-bit
content hash (oracle/dfdb_oracle.c: orc_materialize_hash_blocks); the GPU result is hashed the same way and must agree in
row count and in every column hash (`--check prefix` restores the cheap bit-exact comparison on a leading row range).
One JSON line per query.  bench.py imports run_config() for its `variants`.
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


def table_path(cfg, rows):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    return os.path.join(base, f"dfdb_b200_cfg_{rows}_{cfg['seed']:x}")


def ensure_table(cfg, rows, threads):
    from oracle import oracle as O
    path = table_path(cfg, rows)
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
    if not getattr(run_query, "quiet", False):
        print(json.dumps(line), flush=True)
    return line


def frame_columns(D, np, fr):
    """The result frame as the column list oracle.hash_columns takes."""
    cols = []
    for n in fr.names:
        c = fr[n]
        if isinstance(c, D.FlatStringsVector):
            cols.append(("str", c.sizes, np.frombuffer(c.data, dtype=np.uint8) if not isinstance(c.data, np.ndarray) else c.data))
        elif isinstance(c, np.ma.MaskedArray):
            m = np.ma.getmaskarray(c)
            cols.append((np.where(m, np.zeros((), dtype=c.dtype), c.data), m))
        else:
            cols.append(c)
    return cols


def run_config(config, rows=0, reps=3, check="full", check_rows=2_000_000, quiet=False):
    """Runs the queries of BASELINE.json configs[2] (config=3) / configs[3] (config=4); returns the JSON-able lines."""
    import numpy as np
    import torch
    import dfdb_b200 as D
    from dfdb_b200 import R, _capi
    from oracle import oracle as O

    cfg = CONFIGS[config]
    rows = rows or cfg["rows"]
    threads = os.cpu_count() or 1
    path, info = ensure_table(cfg, rows, threads)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    _capi.init(int(os.environ.get("LOCAL_RANK", "0")))
    L = _capi.lib()
    t = D.open_table(path, mode=D.LOAD_HBM, device=int(os.environ.get("LOCAL_RANK", "0")))
    ot = O.OracleTable(path)
    nchk = min(rows, check_rows)
    nblocks = t.nblocks()

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
            return {"ok": ok, "kind": "prefix", "prefix_rows": nchk, "prefix_selected": got.nrow()}
        return chk

    def check_full(v, fr):
        """Whole table: row count and a content hash of every output column, oracle (all host threads) against the GPU result."""
        t0 = time.time()
        exp, nexp = ot.materialize_hash_mt(D.plan_bytes(v), nblocks, threads, len(fr.names))
        t1 = time.time()
        got = O.hash_columns(frame_columns(D, np, fr), threads)
        ok = nexp == fr.nrow() and [tuple(x) for x in exp] == [tuple(x) for x in got]
        return {"ok": bool(ok), "kind": "full", "rows": rows, "selected_oracle": nexp, "selected_gpu": fr.nrow(),
                "column_hashes": [f"{a:016x}" for a, _ in got], "oracle_s": round(t1 - t0, 1), "oracle_threads": threads}

    def checker(make_prefixed):
        return check_full if check == "full" else check_prefix(make_prefixed)

    if not quiet:
        log(f"[cfg] config {config}: {rows} rows, {info['uncompressed'] / 1e9:.2f} GB decoded, {info['compressed'] / 1e9:.2f} GB compressed")
    lines = []
    if config == 3:
        lines.append(run_query(D, L, 's .== "sony" -> materialize [s]', lambda: t[t.s == "sony", ["s"]], rows, reps,
                               checker(lambda: t[R(1, nchk), :][t.s == "sony", ["s"]])))
        lines.append(run_query(D, L, 'startswith.(s, "s") -> materialize [s, k]', lambda: t[D.startswith(t.s, "s"), ["s", "k"]], rows, reps,
                               checker(lambda: t[R(1, nchk), :][D.startswith(t.s, "s"), ["s", "k"]])))
    else:
        names = [m.name for m in t.meta]
        a, b = names[0], names[1]
        pred = lambda: D.coalesce(getattr(t, a) > 50, False) & D.coalesce(getattr(t, b) < 0.5, False)   # noqa: E731
        lines.append(run_query(D, L, "coalesce.(a .> 50, false) .& coalesce.(b .< 0.5, false) -> materialize [a, b, c, d]",
                               lambda: t[pred(), names], rows, reps, checker(lambda: t[R(1, nchk), :][pred(), names])))
    t.close()
    ot.close()
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3, choices=[3, 4])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", default="full", choices=["full", "prefix"])
    ap.add_argument("--check-rows", type=int, default=2_000_000)
    args = ap.parse_args()
    run_config(args.config, args.rows, args.reps, args.check, args.check_rows)
    return 0


if __name__ == "__main__":
    sys.exit(main())
