"""Oracle pin for the f4 operators: oracle.window / oracle.expand against answers transcribed from the reference's golden files
(tests/window_goldens.py: window.sql.out, group-analytics.sql.out)."""
import math

import pyarrow as pa
import pytest

from oracle import oracle as O
from window_goldens import EXPAND_CASES, EXPAND_TEST_DATA, WINDOW_CASES, WINDOW_CASE_COLUMNS, WINDOW_TEST_DATA


def rows_of(t):
    return list(zip(*[t.column(i).to_pylist() for i in range(t.num_columns)]))


def same_rows(got, want):
    key = lambda r: tuple((x is None, 0 if x is None else x) for x in r)
    g, w = sorted(got, key=key), sorted(want, key=key)
    assert len(g) == len(w)
    for a, b in zip(g, w):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            if isinstance(y, float) and x is not None:
                assert math.isclose(x, y, rel_tol=1e-12), (a, b)
            else:
                assert x == y, (a, b)


@pytest.mark.parametrize("case", WINDOW_CASES, ids=lambda c: "window.sql.out:%d" % c[0])
def test_window_oracle_matches_the_reference_golden(case):
    line, part, orders, specs, want = case
    got = O.window(WINDOW_TEST_DATA, part, orders, specs)
    same_rows(rows_of(got.select(WINDOW_CASE_COLUMNS.get(line, ["val", "cate"]) + [s[4] for s in specs])), want)


def expand_plan(kind, g0, g1, x):
    """GROUP BY g0, g1 WITH CUBE / ROLLUP as the planner lowers it (Analyzer ResolveGroupingAnalytics -> Expand + Aggregate):
    one projection list per grouping set = (agg input, g0 or NULL, g1 or NULL, spark_grouping_id)."""
    ex = {"a": ("col", "a"), "b": ("col", "b"), "a+b": ("add", ("col", "a"), ("col", "b")), "a-b": ("sub", ("col", "a"), ("col", "b"))}
    sets = [(True, True, 0), (True, False, 1), (False, True, 2), (False, False, 3)] if kind == "cube" else [(True, True, 0), (True, False, 1), (False, False, 3)]
    null = ("lit", None)
    return [[ex[x], ex[g0] if k0 else null, ex[g1] if k1 else null, ("lit", gid)] for k0, k1, gid in sets]


@pytest.mark.parametrize("case", EXPAND_CASES, ids=lambda c: "group-analytics.sql.out:%d" % c[0])
def test_expand_then_aggregate_matches_the_reference_golden(case):
    _, kind, g0, g1, x, want = case
    exp = O.expand(EXPAND_TEST_DATA, expand_plan(kind, g0, g1, x), ["x", "g0", "g1", "gid"])
    nsets = 4 if kind == "cube" else 3
    assert exp.num_rows == nsets * EXPAND_TEST_DATA.num_rows
    assert exp.column("gid").to_pylist()[:nsets] == ([0, 1, 2, 3] if kind == "cube" else [0, 1, 3])   # list 0 first, per input row
    agg = O.hash_aggregate(exp, ["g0", "g1", "gid"], [("sum", "x", "s")])
    same_rows(rows_of(agg.select(["g0", "g1", "s"])), want)


def test_window_oracle_frames_by_hand():
    t = pa.table({"p": pa.array([1, 1, 1, 1, 2, 2], type=pa.int32()), "o": pa.array([1, 2, 2, 3, 5, 6], type=pa.int32()),
                  "v": pa.array([10, None, 30, 40, 1, 2], type=pa.int64())})
    got = O.window(t, ["p"], [("o", True, True)],
                   [("sum", "v", ("rows", -1, 1), 0, "s"), ("lag", "v", None, 1, "lag"), ("lead", "v", None, 2, "lead"),
                    ("min", "v", ("rows", None, 0), 0, "m"), ("count", "v", ("range", None, 0), 0, "c")])
    assert got.column("s").to_pylist() == [10, 40, 70, 70, 3, 3]
    assert got.column("lag").to_pylist() == [None, 10, None, 30, None, 1]
    assert got.column("lead").to_pylist() == [30, 40, None, None, None, None]
    assert got.column("m").to_pylist() == [10, 10, 10, 10, 1, 1]
    assert got.column("c").to_pylist() == [1, 2, 2, 3, 1, 2]      # RANGE ... CURRENT ROW includes the peers (o = 2 twice)
