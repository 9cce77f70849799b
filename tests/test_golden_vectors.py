"""Golden fixtures (tests/golden/, written by tests/golden/make_golden.py): tables in the reference's on-disk format whose
blocks were compressed by the system liblz4 -- the reference's own codec -- with expected results computed from the raw
column data by plain Python.  The oracle is checked against them on the CPU; the CUDA path is checked against the same
files and the same expectations on the GPU (through the C ABI).  Neither the oracle nor the CUDA path produced any
expectation in these files."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

import dfdb_b200 as D
from dfdb_b200 import R, _capi
from engines import _norm_col

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
QUERIES = json.load(open(os.path.join(GOLD, "queries.json")))
VECTORS = json.load(open(os.path.join(GOLD, "lz4_vectors.json")))
TABLES = json.load(open(os.path.join(GOLD, "tables.json")))


def _table_path(name):
    return os.path.join(GOLD, "tables", name)


def _same(got, exp):
    """bit-exact for integers / strings / missings and for Float64 values (materialized values are copies)"""
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert (g is None) == (e is None) and (g is None or g == e), (g, e)


def _agg_column(q, t, v):
    """the column a query's `aggregate` expectation is about: `b`, or a computed one"""
    expr = q["aggregate"].get("expr")
    if expr == "a*2+1":
        return v.a * 2 + 1
    if expr == "a*b":
        w = t[t.a > 50, :]
        return w.a * w.b
    return v.b


def _check_aggregate(a, e, sum_f):
    assert a.count == e["count"]
    if "sum_i" in e:
        assert a.sum_i64 == e["sum_i"] and a.min_i64 == e["min_i"] and a.max_i64 == e["max_i"]
        return
    assert abs(sum_f - e["sum"]) <= 1e-12 * abs(e["sum"])                                   # BASELINE.json: 1e-12 relative
    if "min" in e:
        assert a.min_f64 == e["min"] and a.max_f64 == e["max"]


def _rebuild_view(q, t):
    """the query again through the host-side plan algebra: pins the plan serialisation to the committed bytes"""
    n = q["name"]
    if n == "range_predicate_aggregate":
        return t[(t.a > 25) & (t.a <= 75), ["b"]]
    if n == "docs_example":
        return t[t.a > 50, ["b"]]
    if n == "computed_int_aggregate":
        return t[t.a > 50, ["a"]]
    if n == "computed_float_aggregate":
        return t[t.a > 50, ["b"]]
    if n == "string_equality":
        return t[t.s == "sony", ["s", "a"]]
    if n == "string_prefix":
        return t[D.startswith(t.s, "s"), ["s"]]
    if n == "missing_predicate":
        return t[D.coalesce(t.ma > 50, False) & D.coalesce(t.mb < 0.5, False), ["ma", "mb", "sm", "b"]]
    if n == "ismissing_strings":
        return t[D.ismissing(t.sm), ["a", "sm"]]
    if n == "range_after_predicate":
        return t[R(101, 1900), :][t.a > 50, :][R(10, 3, 300), ["a", "s"]]
    if n == "view_jl_mod50":
        return D.selection(D.selection(D.DFView(t), t.a % 50 == 0), t.c < 930)
    if n == "missings_full":
        return D.DFView(t)
    if n == "missings_gt2":
        return t[D.coalesce(t.x > 2, False), ["x"]]
    raise KeyError(n)


# ---- CPU: the oracle against the golden files ----------------------------------------------------------------

@pytest.mark.parametrize("vec", VECTORS, ids=lambda v: f"{v['name']}-accel{v['accel']}")
def test_oracle_codec_decodes_liblz4_streams(oracle, vec):
    body = oracle.lz4_decompress(base64.b64decode(vec["compressed"]), vec["origin"])
    assert len(body) == vec["origin"] and hashlib.sha256(body).hexdigest() == vec["sha256"]


@pytest.mark.parametrize("q", QUERIES, ids=lambda q: q["name"])
def test_oracle_matches_golden_queries(oracle, q):
    t = D.open_table(_table_path(q["table"]))
    ot = oracle.OracleTable(_table_path(q["table"]))
    try:
        v = _rebuild_view(q, t)
        plan = D.plan_bytes(v)
        assert plan.hex() == q["plan"], "the plan algebra no longer emits the committed plan bytes"
        assert ot.count(plan) == q["count"]
        mask = ot.mask(plan)
        assert (np.nonzero(mask)[0] + 1).tolist() == q["rows"]
        cols = ot.materialize(plan)
        for name, c in zip(v.projection.keys(), cols):
            _same(_norm_col(c), q["columns"][name])
        if "aggregate" in q:
            a = ot.aggregate(D.plan_bytes(_agg_column(q, t, v)), 0)
            _check_aggregate(a, q["aggregate"], a.sum_kahan)
    finally:
        ot.close()
        t.close()


def test_golden_tables_hold_the_raw_data(oracle):
    """the files themselves: every column decodes (oracle reader) to the raw data the expectations were computed from"""
    for name, tj in TABLES.items():
        t = D.open_table(_table_path(name))
        ot = oracle.OracleTable(_table_path(name))
        try:
            v = D.DFView(t)
            cols = ot.materialize(D.plan_bytes(v))
            for cname, c in zip(v.projection.keys(), cols):
                _same(_norm_col(c), tj["columns"][cname]["data"])
        finally:
            ot.close()
            t.close()


# ---- GPU: the CUDA path against the same files ----------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("flavour", [2, 4, 5, 6], ids=["walker_consumer", "verified_runs", "long_sequences", "byte_streams"])
def test_gpu_codec_decodes_liblz4_streams(flavour):
    _capi.init(0)
    L = _capi.lib()
    blocks = [base64.b64decode(v["compressed"]) for v in VECTORS]
    n = len(blocks)
    comp = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    clen = np.array([len(b) for b in blocks], dtype=np.int64)
    coff = np.concatenate([[0], np.cumsum(clen)[:-1]]).astype(np.int64)
    orig = np.array([v["origin"] for v in VECTORS], dtype=np.int64)
    ooff = np.concatenate([[0], np.cumsum(orig)[:-1]]).astype(np.int64)
    out = np.zeros(max(int(orig.sum()), 1), dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    _capi.check(L.dfdb_set_option(b"lz4_flavour", flavour))
    try:
        _capi.check(L.dfdb_lz4_decode_blocks(comp.ctypes.data, coff.ctypes.data, clen.ctypes.data, out.ctypes.data, ooff.ctypes.data,
                                             orig.ctypes.data, n, status.ctypes.data))
    finally:
        L.dfdb_set_option(b"lz4_flavour", 0)
    for k, v in enumerate(VECTORS):
        assert status[k] == 0, (v["name"], status[k])
        assert hashlib.sha256(bytes(out[ooff[k]:ooff[k] + orig[k]])).hexdigest() == v["sha256"], v["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("q", QUERIES, ids=lambda q: q["name"])
def test_gpu_matches_golden_queries(q):
    _capi.init(0)
    t = D.open_table(_table_path(q["table"]))
    try:
        v = _rebuild_view(q, t)
        assert D.nrow(v) == q["count"]
        assert (np.nonzero(D.selection_mask(v))[0] + 1).tolist() == q["rows"]
        assert D.selection_indices(v).tolist() == q["rows"]
        fr = D.materialize(v).to_dict()
        for name in v.projection.keys():
            _same(fr[name], q["columns"][name])
        if "aggregate" in q:
            a = D.aggregate(_agg_column(q, t, v))
            _check_aggregate(a, q["aggregate"], a.sum_f64 + a.sum_f64_lo)
    finally:
        t.close()
