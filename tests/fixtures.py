"""Fixture tables written in the reference's on-disk format (see oracle/oracle.py write_table)."""
import numpy as np

from cases import MISSING_VEC, STR_MISSING_VEC, STR_VEC


def make_reference_fixture(O, path, sz=1000, block_size=100):
    """test/view.jl:8-15 : a = 1:sz, b = string.(1:sz), c = 1:sz ; block_size = 100 (10 blocks)."""
    a = np.arange(1, sz + 1, dtype=np.int64)
    b = [str(i) for i in range(1, sz + 1)]
    c = np.arange(1, sz + 1, dtype=np.int64)
    O.write_table(path, [("a", "Int64", a), ("b", "String", b), ("c", "Int64", c)], block_size=block_size)
    return {"a": a, "b": b, "c": c}


def make_selection_fixture(O, path, block_size):
    """test/selection.jl:46,74-77 : a = 1:100, b = a * 5"""
    a = np.arange(1, 101, dtype=np.int64)
    O.write_table(path, [("a", "Int64", a), ("b", "Int64", a * 5)], block_size=block_size)
    return {"a": a, "b": a * 5}


def make_broadcast_fixture(O, path, block_size=32):
    """test/broadcast.jl:6-10 : a = 1:100, b = string.(1:100), c = 0.5:0.5:50"""
    a = np.arange(1, 101, dtype=np.int64)
    c = np.arange(1, 101, dtype=np.float64) * 0.5
    O.write_table(path, [("a", "Int64", a), ("b", "String", [str(i) for i in a]), ("c", "Float64", c)], block_size=block_size)
    return {"a": a, "c": c}


def make_missing_fixture(O, path, block_size=4):
    """test/missings.jl:4 : [1, missing, 2, 3, missing, 5, 6, missing, 10, 11, missing] as Union{Int64,Missing}"""
    x = np.array([0 if v is None else v for v in MISSING_VEC], dtype=np.int64)
    m = np.array([v is None for v in MISSING_VEC])
    O.write_table(path, [("x", "Missing(Int64)", (x, m))], block_size=block_size)
    return {}


def make_strings_fixture(O, path, block_size=4):
    """test/flat_strings.jl:65,80"""
    O.write_table(path, [("s", "String", STR_VEC), ("sm", "Missing(String)", STR_MISSING_VEC)], block_size=block_size)
    return {}
