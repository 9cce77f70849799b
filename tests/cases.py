"""Known-answer cases of the reference's own tests for the scan path, restated against independent
numpy / Python expectations (the reference tests compare with an in-memory DataFrame the same way).

Every case takes an engine `E` (oracle or GPU, see engines.py), a table opened through the product's
host API (plans are built with the product's plan algebra either way) and the raw fixture data.
Citations are /root/reference/test/<file>:<lines>.
"""
import numpy as np
import pytest

import dfdb_b200 as D
from dfdb_b200 import R


def _rows(data, idx):
    """expected full-row materialisation for 0-based row indices"""
    return {"a": data["a"][idx].tolist(), "b": [data["b"][i] for i in idx], "c": data["c"][idx].tolist()}


# ---- test/view.jl -------------------------------------------------------------------------------

def case_view_full(E, t, data):
    # view.jl:30-39  materialize(v1) == df ; nrow ; size
    v1 = D.DFView(t)
    assert E.materialize(v1) == _rows(data, np.arange(1000))
    assert E.nrow(v1) == 1000
    v2 = D.selection(v1, R(1, 1000))
    assert E.materialize(v2) == _rows(data, np.arange(1000))
    assert E.nrow(v2) == 1000


def case_view_predicates(E, t, data):
    # view.jl:19-50  a % 50 == 0 ; & c < 930 ; projection a / 50
    v1 = D.DFView(t)
    v2 = D.selection(v1, R(1, 1000))
    v3 = D.selection(v2, t.a % 50 == 0)
    idx = np.nonzero(data["a"] % 50 == 0)[0]
    assert E.materialize(v3) == _rows(data, idx)
    assert E.nrow(v3) == len(idx)
    v4 = D.selection(v3, v3.c < 930)
    v4 = D.projection(v4, {"a": v4.a / 50})
    ind = (data["a"] % 50 == 0) & (data["c"] < 930)
    assert E.materialize(v4) == {"a": (data["a"][ind] / 50).tolist()}
    assert E.nrow(v4) == int(ind.sum())
    with pytest.raises(D.ArgumentError):
        D.projection(v4, {"c": "c"})          # view.jl:54 view don't have column :c


def case_view_projections(E, t, data):
    v1 = D.DFView(t)
    # view.jl:56-70
    assert E.materialize(D.projection(v1, ["a", "c"])) == {"a": data["a"].tolist(), "c": data["c"].tolist()}
    tv = D.projection(v1, {"a": "a", "c": v1.c * 2})
    assert E.materialize(tv) == {"a": data["a"].tolist(), "c": (data["c"] * 2).tolist()}
    assert E.materialize(D.projection(v1, [1, 3])) == {"a": data["a"].tolist(), "c": data["c"].tolist()}
    assert E.materialize(D.projection(v1, R(1, 2))) == {"a": data["a"].tolist(), "b": list(data["b"])}
    # view.jl:72-93
    idx = np.nonzero(data["a"] % 50 == 0)[0]
    assert E.materialize(D.selproj(v1, t.a % 50 == 0, ["c"])) == {"c": data["c"][idx].tolist()}
    assert E.materialize(D.selproj(v1, 1, ["c"])) == {"c": [1]}
    assert E.materialize(D.selproj(v1, [1, 200], ["c"])) == {"c": [1, 200]}
    assert E.materialize(v1[[1, 200], ["c"]]) == {"c": [1, 200]}
    # view.jl:95-122
    tv = v1[R(1, 200), :]
    assert E.nrow(tv) == 200 and E.materialize(tv) == _rows(data, np.arange(200))
    tv = v1[:, {"e": "a"}]
    assert E.nrow(tv) == 1000 and E.materialize(tv) == {"e": data["a"].tolist()}
    assert v1[:, :] is v1
    tv = t[R(1000 - 10, 1000), {"e": "a"}]          # end-10:end
    assert E.nrow(tv) == 11 and E.materialize(tv) == {"e": data["a"][-11:].tolist()}
    tv = t[R(1000 - 10, 1000), R(3 - 1, 3)]         # end-10:end, end-1:end
    assert E.materialize(tv) == {"b": list(data["b"][-11:]), "c": data["c"][-11:].tolist()}
    # view.jl:125-137
    assert t[R(1, 20), ["a", "b"]] == t[R(1, 20), ["a", "b"]]
    assert t[R(1, 30), ["a", "b"]] != t[R(1, 20), ["a", "b"]]
    assert not D.issameselection(t[R(1, 30), ["a", "b"]], t[R(1, 20), ["a", "b"]])
    assert t[R(1, 20), ["a", "b"]] != t[R(1, 20), ["b", "a"]]


# ---- test/range_indexing.jl ------------------------------------------------------------------------

def case_range_indexing(E, t, data):
    # range_indexing.jl:13-27
    assert E.materialize(t[:, ["a"]]) == {"a": data["a"].tolist()}
    assert E.materialize(t[R(5, 60), :]) == _rows(data, np.arange(4, 60))
    assert E.materialize(t[R(5, 300), :]) == _rows(data, np.arange(4, 300))
    assert E.materialize(t[R(5, 300, 1000), :]) == _rows(data, np.arange(4, 1000, 300))
    assert E.materialize(t[[1, 200, 20], :]) == _rows(data, np.array([0, 19, 199]))     # table order
    assert E.nrow(t[R(5, 60), :]) == 56
    assert E.nrow(t[R(5, 2, 60), :]) == len(range(5, 61, 2))
    assert E.materialize(t[R(1000 - 20, 1000), :]) == _rows(data, np.arange(979, 1000))


def case_range_composition(E, t, data):
    # docs/src/index.md:127-145: v[1:10:end][1:10] == 1:10:91 ; selection.jl:16-22
    v = t[R(1, 10, 1000), :][R(1, 10), :]
    assert len(v.selection) == 1 and v.selection.queue[0] == R(1, 10, 91)
    assert E.materialize(v[:, ["a"]]) == {"a": list(range(1, 92, 10))}
    # range after predicate counts survivors across blocks (selection.jl:94-111)
    v = t[t.a % 7 == 0, :][R(3, 2, 40), ["a"]]
    surv = data["a"][data["a"] % 7 == 0]
    assert E.materialize(v) == {"a": surv[2:40:2].tolist()}
    # index vector after predicate, then another predicate
    v = t[t.a > 100, :][[5, 1, 300, 5], :][t.c < 500, ["c"]]
    surv = data["c"][data["a"] > 100]
    pick = surv[[0, 4, 299]]
    assert E.materialize(v) == {"c": pick[pick < 500].tolist()}


# ---- test/selection.jl -----------------------------------------------------------------------------

def case_selection_stages(E, t100, data100):
    """t100: a = 1:100, b = a*5 ; used with block_size 100 (one block) and 50 (two blocks)."""
    t = t100
    # selection.jl:40-49  5:20 |> 3:4 => rows 7:8
    v = t[R(5, 20), :][R(3, 4), ["a"]]
    assert E.materialize(v) == {"a": [7, 8]}
    # selection.jl:51-72  10:60 |> (65 > a > 34) |> 15:18  => rows 49:52 (also across two 50-row blocks)
    v = t[R(10, 60), :]
    v = v[(v.a < 65) & (v.a > 34), :]
    v = v[R(15, 18), ["a"]]
    assert len(v.selection) == 3
    assert E.materialize(v) == {"a": [49, 50, 51, 52]}
    assert E.nrow(v) == 4
    # selection.jl:74-106 two predicates fuse into one stage
    v = t[(t.a < 65) & (t.a > 34), :]
    v = v[v.b % 10 == 0, :]
    assert len(v.selection) == 1
    a, b = data100["a"], data100["b"]
    ind = (65 > a) & (a > 34) & (b % 10 == 0)
    assert E.materialize(v) == {"a": a[ind].tolist(), "b": b[ind].tolist()}
    # selection.jl:36-37 non-Bool predicate
    with pytest.raises(D.ArgumentError):
        t[t.a * 3, :]


# ---- test/broadcast.jl, test/projection.jl ------------------------------------------------------------

def case_broadcast_eval(E, tb, datab):
    """tb: a = 1:100 (Int64), b = string.(1:100), c = 0.5:0.5:50 (Float64)."""
    a, c = datab["a"], datab["c"]
    # broadcast.jl:46-52 gather on a stride-10 selection then a + (a + c)  (== 2a + c)
    v = tb[R(1, 10, 100), {"x": tb.a + (tb.a + tb.c)}]
    assert E.materialize(v) == {"x": (a[0:100:10] * 2 + c[0:100:10]).tolist()}
    # broadcast.jl:54-61 scalar argument
    v = tb[R(1, 10, 100), {"x": tb.a + 20}]
    assert E.materialize(v) == {"x": (a[0:100:10] + 20).tolist()}
    # broadcast.jl:63-71 in.(a, Ref([1, 11, 21]))
    v = tb[R(1, 10, 100), {"x": D.isin(tb.a, [1, 11, 21])}]
    assert E.materialize(v) == {"x": np.isin(a[0:100:10], [1, 11, 21]).tolist()}
    # broadcast.jl:73-81 arrays are rejected
    with pytest.raises(D.ArgumentError):
        D.isin(tb.a, [1, 11, 21]) & np.array([1, 2, 3])
    # projection.jl:57-80 plain + computed column on 1:10:100
    v = tb[R(1, 10, 100), {"a": "a", "b": tb.a * 2}]
    assert E.materialize(v) == {"a": a[0:100:10].tolist(), "b": (a[0:100:10] * 2).tolist()}
    # eltype inference broadcast.jl:22-36
    assert (tb.a * 2).eltype() == D.JType("Int64")
    assert (tb.a + tb.c).eltype() == D.JType("Float64")
    assert (tb.a + (tb.a + tb.c)).eltype() == D.JType("Float64")


# ---- test/columnbroadcast.jl ---------------------------------------------------------------------------

def case_column_broadcast(E, t, data):
    a, b, c = data["a"], data["b"], data["c"]
    # columnbroadcast.jl:26 different selections
    with pytest.raises(D.ArgumentError):
        t.a[R(1, 20)] + t.c[R(11, 30)]
    # columnbroadcast.jl:28-33
    assert E.column(t.a[R(1, 20)] + 20) == (a[:20] + 20).tolist()
    assert E.column(t.a[R(1, 20)] * t.a[R(1, 20)]) == (a[:20] * a[:20]).tolist()
    assert E.column(t.a[R(1, 20)] * t.a[R(1, 20)] - 20) == (a[:20] * a[:20] - 20).tolist()
    assert E.column(t.a * t.c) == (a * c).tolist()
    assert E.column(t.a == 10) == (a == 10).tolist()
    # columnbroadcast.jl:45-53  300 .>= a .>= 10 then startswith.(b, "1")
    tb2 = t[(t.a <= 300) & (t.a >= 10), :]
    i2 = np.nonzero((a <= 300) & (a >= 10))[0]
    assert E.materialize(tb2) == _rows(data, i2)
    tb3 = tb2[D.startswith(tb2.b, "1"), :]
    i3 = np.array([i for i in i2 if b[i].startswith("1")])
    assert E.materialize(tb3) == _rows(data, i3)
    # columnbroadcast.jl:55-60 DFView from columns
    v = D.view_from_columns(a=t.a * 3, g=t.a * t.c)
    assert E.materialize(v) == {"a": (a * 3).tolist(), "g": (a * c).tolist()}


# ---- test/column.jl ----------------------------------------------------------------------------------------

def case_columns(E, t, data):
    col = t[:, 1]
    assert isinstance(col, D.DFColumn) and isinstance(t[:, [1]], D.DFView) and isinstance(t[:, ["a"]], D.DFView)
    assert isinstance(t[:, "a"], D.DFColumn) and isinstance(t[R(1, 5, 1000), "a"], D.DFColumn)
    assert E.nrow(col.view) == 1000                                   # column.jl:29
    assert E.column(col) == data["a"].tolist()                        # column.jl:31
    assert col.eltype() == D.JType("Int64")
    col2 = col[R(90, 110)]
    assert E.column(col2) == data["a"][89:110].tolist()               # column.jl:37
    assert E.column(col2[R(1, 1)]) == [90]                            # col2[1] == 90
    assert E.column(D.DFColumn(D.selection(col2.view, 12))) == [101]  # col2[12] == 101 (column.jl:40)
    assert E.column(t.a + t.c * 2) == (data["a"] + data["c"] * 2).tolist()   # column.jl:42-43
    assert E.column(t.a * 4) == (data["a"] * 4).tolist()
    assert t.a.same_as(t[:, "a"]) and t[R(1, 20), :].a.same_as(t[R(1, 20), "a"])


# ---- test/missings.jl, test/flat_strings.jl ------------------------------------------------------------------

MISSING_VEC = [1, None, 2, 3, None, 5, 6, None, 10, 11, None]            # missings.jl:4
STR_VEC = ["1", "222", "32", "44", "335", "11116", "312313127", "444", "assadf", "bvxvbx"]       # flat_strings.jl:65
STR_MISSING_VEC = ["1", "222", None, "44", "335", "11116", None, "444", "assadf", "bvxvbx"]      # flat_strings.jl:80


def case_missings(E, tm, _):
    """tm: x = MISSING_VEC as Union{Int64,Missing}."""
    assert E.materialize(tm[:, ["x"]]) == {"x": MISSING_VEC}                                     # missings.jl:1-11
    assert E.materialize(tm[D.ismissing(tm.x), ["x"]]) == {"x": [None] * 4}
    assert E.materialize(tm[~D.ismissing(tm.x), ["x"]]) == {"x": [v for v in MISSING_VEC if v is not None]}
    assert E.materialize(tm[D.coalesce(tm.x > 2, False), ["x"]]) == {"x": [3, 5, 6, 10, 11]}
    assert E.materialize(tm[D.coalesce(tm.x > 2, True), ["x"]]) == {"x": [None, 3, None, 5, 6, None, 10, 11, None]}
    with pytest.raises(D.ArgumentError):          # Union{Missing,Bool} predicate (selection.jl:52-55)
        tm[tm.x > 2, :]
    # three-valued logic: (missing & false) == false, (missing | true) == true
    v = tm[D.coalesce((tm.x > 2) & (tm.x < 0), False), ["x"]]
    assert E.materialize(v) == {"x": []}
    v = tm[D.coalesce((tm.x > 2) | D.ismissing(tm.x), False), ["x"]]
    assert E.materialize(v) == {"x": [None, 3, None, 5, 6, None, 10, 11, None]}


def case_flat_strings(E, tm, _):
    # flat_strings.jl:64-78 gathers by 3:5, 1:2:10, [:], startswith mask
    assert E.materialize(tm[R(3, 5), ["s"]]) == {"s": STR_VEC[2:5]}
    assert E.materialize(tm[R(1, 2, 10), ["s"]]) == {"s": STR_VEC[0:10:2]}
    assert E.materialize(tm[R(1, 10), ["s"]]) == {"s": STR_VEC}
    assert E.materialize(tm[D.startswith(tm.s, "3"), ["s"]]) == {"s": [s for s in STR_VEC if s.startswith("3")]}
    assert E.materialize(tm[D.endswith(tm.s, "4"), ["s"]]) == {"s": [s for s in STR_VEC if s.endswith("4")]}
    assert E.materialize(tm[tm.s == "444", ["s"]]) == {"s": ["444"]}
    assert E.materialize(tm[tm.s != "444", ["s"]]) == {"s": [s for s in STR_VEC if s != "444"]}
    assert E.materialize(tm[tm.s < "4", ["s"]]) == {"s": [s for s in STR_VEC if s < "4"]}
    # flat_strings.jl:80-97 with missing: ranges keep missing, ismissing masks
    assert E.materialize(tm[R(3, 5), ["sm"]]) == {"sm": STR_MISSING_VEC[2:5]}
    assert E.materialize(tm[R(1, 2, 10), ["sm"]]) == {"sm": STR_MISSING_VEC[0:10:2]}
    assert E.materialize(tm[D.ismissing(tm.sm), ["sm"]]) == {"sm": [None, None]}
    assert E.materialize(tm[~D.ismissing(tm.sm), ["sm"]]) == {"sm": [s for s in STR_MISSING_VEC if s is not None]}
    assert E.materialize(tm[D.coalesce(tm.sm == "444", False), ["sm", "s"]]) == {"sm": ["444"], "s": ["444"]}
    # empty selections: FlatStringsVector a[[]] has length 0 (flat_strings.jl:88); an empty *row* selection
    # throws when the RangeToProcess is built (minimum of an empty collection, selection.jl:73)
    with pytest.raises(D.ArgumentError):
        E.materialize(tm[R(1, 0), ["s"]])
    assert E.materialize(tm[tm.s == "no such", ["s", "sm"]]) == {"s": [], "sm": []}


# ---- aggregates (column.jl:102-126 + Base folds) -------------------------------------------------------------

def case_aggregates(E, t, data):
    a = data["a"]
    ag = E.agg(t.a)
    assert (ag.count, ag.nmissing, ag.sum_i64, ag.min_i64, ag.max_i64) == (1000, 0, int(a.sum()), 1, 1000)
    sel = t[(t.a > 25) & (t.a <= 75), :]
    ag = E.agg(sel.c)
    assert (ag.count, ag.sum_i64, ag.min_i64, ag.max_i64) == (50, int(a[25:75].sum()), 26, 75)
    ag = E.agg(t[t.a > 5000, :].a)
    assert ag.count == 0
    # range |> predicate |> range, then aggregate
    v = t[R(10, 600), :]
    v = v[v.a % 3 == 0, :][R(2, 50), :]
    exp = a[9:600][a[9:600] % 3 == 0][1:50]
    ag = E.agg(v.c)
    assert (ag.count, ag.sum_i64, ag.min_i64, ag.max_i64) == (len(exp), int(exp.sum()), int(exp.min()), int(exp.max()))
