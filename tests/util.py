"""Shared helpers for the parity tests: fixture tables in the reference's on-disk format and
plan construction with the product's plan algebra (pure Python, no GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dfdb_b200 as D  # noqa: E402  (loader for the `dataframedbs.jl_b200/` package)
from dfdb_b200.plan import (BlockBroadcasting as BB, ColRef, JType, Projection, R, SelectionQueue, add,  # noqa: E402,F401
                            encode_plan, InSet)


def reference_fixture(path, oracle, sz=1000, block_size=100):
    """test/view.jl:8-15 : a = 1:sz, b = string.(1:sz), c = 1:sz ; block_size = 100."""
    a = np.arange(1, sz + 1, dtype=np.int64)
    b = [str(i) for i in range(1, sz + 1)]
    c = np.arange(1, sz + 1, dtype=np.int64)
    oracle.write_table(path, [("a", "Int64", a), ("b", "String", b), ("c", "Int64", c)], block_size=block_size)
    return {"a": a, "b": b, "c": c}


def colrefs(columns):
    """[(id, name, typestring)] -> {name: ColRef}"""
    return {name: ColRef(name, JType.parse(ts), cid) for cid, name, ts in columns}


def full_projection(columns):
    refs = colrefs(columns)
    return Projection([(name, refs[name]) for _, name, _ in columns])


def make_plan(stages, proj):
    q = SelectionQueue()
    for s in stages:
        q = add(q, s)
    return encode_plan(q, proj)
