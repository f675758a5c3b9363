"""Host-side logic that needs no GPU: query parsing (reference query.c:44-249
semantics) and the in-memory relation_map."""
import numpy as np
import pytest


def test_parse_query_orders_predicates_like_the_reference(b200):
    q = b200.parse_query("3 0 1|0.2=1.0&0.1=2.0&0.2>3499&1.1<7|1.2 0.1\n")
    assert q.relations == [3, 0, 1]
    # filters go to the list head => reverse textual order (query.c:150-157)
    assert q.filters == [(1, 1, "<", 7), (0, 2, ">", 3499)]
    assert q.joins == [(0, 2, 1, 0), (0, 1, 2, 0)]
    assert q.views == [(1, 2), (0, 1)]


def test_parse_query_rejects_what_the_reference_cannot_run(b200):
    with pytest.raises(ValueError):
        b200.parse_query("0 1|3<0.1&0.0=1.0|0.0")      # constant on the left (quirk 3)
    with pytest.raises(ValueError):
        b200.parse_query("0 1|0.0<1.0|0.0")            # non-equi join


def test_relation_map_layout(b200):
    cols = [[np.arange(10, dtype=np.uint64), np.arange(10, dtype=np.uint64) * 3], [np.array([7, 7, 9], np.uint64)]]
    rm = b200.RelationMapArray(cols)
    assert len(rm) == 2
    assert rm.array[0].num_tuples == 10 and rm.array[0].num_columns == 2
    assert rm.array[0].columns[1][4] == 12
    assert rm.array[1].col_stats[0].l == 7 and rm.array[1].col_stats[0].u == 9 and rm.array[1].col_stats[0].d == 2.0


def test_null_line(b200):
    assert b200.QueryResult(None, 3).line() == "NULL NULL NULL"
    assert b200.QueryResult([1, 2], 5).line() == "1 2"


def test_scripts_and_entry_points_compile():
    """Every Python entry point of the repo (bench, graft entry, scripts, the multi-GPU worker) byte-compiles and
    every shell script parses: they only run on GPU boxes, where a syntax error would cost GPU minutes."""
    import py_compile
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    files = [root / "bench.py", root / "__graft_entry__.py", root / "tests" / "multi_gpu_worker.py"]
    files += sorted((root / "scripts").glob("*.py")) + sorted((root / "profiles").glob("*.py"))
    for f in files:
        py_compile.compile(str(f), doraise=True)
    for sh in sorted((root / "scripts").glob("*.sh")):
        assert subprocess.run(["bash", "-n", str(sh)]).returncode == 0, sh
