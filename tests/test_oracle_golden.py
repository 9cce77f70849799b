"""CPU: the oracle against the reference's own known-answer tests (cases.py), plans built by the product's
host-side plan algebra.  Runs without a GPU."""
import os

import pytest

import cases
import dfdb_b200 as D
import fixtures
from engines import OracleEngine


@pytest.fixture(scope="module")
def ref_table(tmp_path_factory, oracle):
    p = str(tmp_path_factory.mktemp("ref") / "test_data")
    data = fixtures.make_reference_fixture(oracle, p)
    t = D.open_table(p)
    E = OracleEngine(oracle, p)
    yield E, t, data
    E.close()
    t.close()


@pytest.mark.parametrize("case", [cases.case_view_full, cases.case_view_predicates, cases.case_view_projections,
                                  cases.case_range_indexing, cases.case_range_composition, cases.case_column_broadcast,
                                  cases.case_columns, cases.case_aggregates], ids=lambda f: f.__name__)
def test_reference_fixture_cases(ref_table, case):
    case(*ref_table)


@pytest.mark.parametrize("block_size", [100, 50, 7])
def test_selection_stages(tmp_path, oracle, block_size):
    p = str(tmp_path / "sel")
    data = fixtures.make_selection_fixture(oracle, p, block_size)
    t = D.open_table(p)
    E = OracleEngine(oracle, p)
    cases.case_selection_stages(E, t, data)


def test_broadcast_eval(tmp_path, oracle):
    p = str(tmp_path / "bc")
    data = fixtures.make_broadcast_fixture(oracle, p)
    cases.case_broadcast_eval(OracleEngine(oracle, p), D.open_table(p), data)


@pytest.mark.parametrize("block_size", [4, 64, 3])
def test_missings(tmp_path, oracle, block_size):
    p = str(tmp_path / "ms")
    fixtures.make_missing_fixture(oracle, p, block_size)
    cases.case_missings(OracleEngine(oracle, p), D.open_table(p), None)


@pytest.mark.parametrize("block_size", [4, 64])
def test_flat_strings(tmp_path, oracle, block_size):
    p = str(tmp_path / "st")
    fixtures.make_strings_fixture(oracle, p, block_size)
    cases.case_flat_strings(OracleEngine(oracle, p), D.open_table(p), None)


def test_open_table_errors(tmp_path, oracle):
    with pytest.raises(RuntimeError):
        D.open_table(str(tmp_path / "nope"))                      # creators.jl:8
    p = str(tmp_path / "t")
    fixtures.make_selection_fixture(oracle, p, 50)
    os.remove(os.path.join(p, "2.bin"))
    with pytest.raises(RuntimeError):
        D.open_table(p)                                           # filesystem.jl:58
    with pytest.raises(oracle.OracleError):
        oracle.OracleTable(p)
