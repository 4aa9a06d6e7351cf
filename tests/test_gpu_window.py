"""GPU parity for the f4 operators (WindowExec, ExpandExec, SortAggregateExec) against the reference's golden answers
(tests/window_goldens.py) and against the oracle on random inputs."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from test_window_cpu import expand_plan, rows_of, same_rows
from util import assert_tables_equal
from window_goldens import EXPAND_CASES, EXPAND_TEST_DATA, WINDOW_CASES, WINDOW_CASE_COLUMNS, WINDOW_TEST_DATA

pytestmark = pytest.mark.gpu


def _window(table, part, orders, specs, stream):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, WindowExec, WindowFunction
    fns = [WindowFunction(f, c, fr, p, name) for f, c, fr, p, name in specs]
    return WindowExec(fns, part, orders, LocalTableScanExec(ColumnarBatch.from_arrow(table, stream))).collect(stream)


@pytest.mark.parametrize("case", WINDOW_CASES, ids=lambda c: "window.sql.out:%d" % c[0])
def test_window_matches_the_reference_golden(gpu, stream, case):
    line, part, orders, specs, want = case
    got = _window(WINDOW_TEST_DATA, part, orders, specs, stream)
    same_rows(rows_of(got.select(WINDOW_CASE_COLUMNS.get(line, ["val", "cate"]) + [s[4] for s in specs])), want)


def _expr(x):
    from spark_b200.expressions import Literal, col
    from spark_b200 import expressions as E
    if x[0] == "col":
        return col(x[1])
    if x[0] == "lit":
        return Literal(x[1], E.VT_I32)
    a, b = _expr(x[1]), _expr(x[2])
    return a + b if x[0] == "add" else a - b


@pytest.mark.parametrize("case", EXPAND_CASES, ids=lambda c: "group-analytics.sql.out:%d" % c[0])
@pytest.mark.parametrize("agg_cls", ["hash", "sort"])
def test_expand_then_aggregate_matches_the_reference_golden(gpu, stream, case, agg_cls):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import ExpandExec, HashAggregateExec, LocalTableScanExec, SortAggregateExec
    from spark_b200.expressions import Sum, col
    _, kind, g0, g1, x, want = case
    plan = expand_plan(kind, g0, g1, x)
    exp = ExpandExec([[_expr(e) for e in p] for p in plan], ["x", "g0", "g1", "gid"], LocalTableScanExec(ColumnarBatch.from_arrow(EXPAND_TEST_DATA, stream)))
    expanded = exp.collect(stream)
    want_exp = O.expand(EXPAND_TEST_DATA, plan, ["x", "g0", "g1", "gid"])
    assert rows_of(expanded) == rows_of(want_exp)                    # the reference's row order: per input row, list 0 first
    cls = HashAggregateExec if agg_cls == "hash" else SortAggregateExec
    got = cls(["g0", "g1", "gid"], [(Sum(col("x")), "s")], exp).collect(stream)
    same_rows(rows_of(got.select(["g0", "g1", "s"])), want)


@pytest.mark.parametrize("n,nparts", [(1, 1), (2049, 3), (100_000, 1), (100_000, 700), (60_000, 60_000)])
def test_window_functions_against_the_oracle(gpu, stream, n, nparts):
    """Every function and frame shape on random data: several scan tiles, partitions that span tiles, single-row partitions,
    NULL inputs, double inputs (1e-6 relative on sums, tests/util.py), string partition keys."""
    rng = np.random.default_rng(n + nparts)
    words = np.array(["", "a", "ab", "b", "zz", "你好"])
    t = pa.table({"p": pa.array(rng.integers(0, nparts, n), type=pa.int64(), mask=rng.random(n) < 0.02),
                  "s": pa.array(words[rng.integers(0, len(words), n)], type=pa.string(), mask=rng.random(n) < 0.05),
                  "o": pa.array(rng.integers(0, 50, n), type=pa.int32(), mask=rng.random(n) < 0.03),
                  "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64(), mask=rng.random(n) < 0.1),
                  "d": pa.array(rng.random(n) * 100, type=pa.float64(), mask=rng.random(n) < 0.1),
                  "row": np.arange(n, dtype=np.int64)})
    specs = [("row_number", None, None, 0, "rn"), ("rank", None, None, 0, "rk"), ("dense_rank", None, None, 0, "dr"),
             ("percent_rank", None, None, 0, "pr"), ("cume_dist", None, None, 0, "cd"), ("ntile", None, None, 3, "nt"),
             ("lag", "v", None, 1, "lag1"), ("lead", "d", None, 2, "lead2"),
             ("sum", "v", None, 0, "sum_def"), ("sum", "d", ("rows", None, 0), 0, "sum_run"), ("count", "v", ("rows", -2, 2), 0, "cnt_slide"),
             ("avg", "v", ("rows", -3, 0), 0, "avg_slide"), ("sum", "v", ("rows", 0, None), 0, "sum_rest"), ("sum", "v", ("rows", None, None), 0, "sum_all"),
             ("min", "v", None, 0, "min_def"), ("max", "d", ("rows", None, 0), 0, "max_run"), ("min", "d", ("rows", None, None), 0, "min_all"),
             ("first_value", "v", None, 0, "fv"), ("last_value", "v", None, 0, "lv"), ("last_value", "d", ("rows", -1, 1), 0, "lv_slide")]
    orders = [("o", False, False), ("row", True, True)]       # `row` makes the order total: one answer for row_number / lag / lead
    got = _window(t, ["p", "s"], orders, specs, stream)
    want = O.window(t, ["p", "s"], orders, specs)
    assert got.column("row").to_pylist() == want.column("row").to_pylist()          # output order = partition keys ++ order spec
    assert_tables_equal(got, want, ordered=True)
    # peers: without the tie-breaker the default RANGE frame takes whole peer groups
    specs2 = [("rank", None, None, 0, "rk"), ("dense_rank", None, None, 0, "dr"), ("cume_dist", None, None, 0, "cd"),
              ("sum", "v", None, 0, "sum_def"), ("count", "d", None, 0, "cnt_def"), ("max", "v", None, 0, "max_def"), ("last_value", "o", None, 0, "lv")]
    got2 = _window(t, ["p"], [("o", True, True)], specs2, stream)
    want2 = O.window(t, ["p"], [("o", True, True)], specs2)
    key = ["p", "o", "row"]
    assert_tables_equal(got2, want2, key_cols=key)
    # RANGE frames with value offsets: integer key ascending and descending (NULL keys keep to their peers), double key
    # (the oracle re-evaluates every sliding frame: only where partitions are small)
    if n // nparts > 1000:
        return
    for orders3, specs3 in (([("o", True, True)], [("sum", "v", ("range", -3, 2), 0, "rs"), ("count", "d", ("range", -1, None), 0, "rc"), ("max", "v", ("range", None, 4), 0, "rm")]),
                            ([("o", False, False)], [("sum", "v", ("range", -3, 2), 0, "rs"), ("avg", "d", ("range", 0, 5), 0, "ra")]),
                            ([("d", True, False)], [("count", "v", ("range", -2.5, 0.75), 0, "rc"), ("sum", "d", ("range", None, 1.5), 0, "rs")])):
        got3 = _window(t, ["p"], orders3, specs3, stream)
        want3 = O.window(t, ["p"], orders3, specs3)
        assert_tables_equal(got3, want3, key_cols=["p", orders3[0][0], "row"])


# sql-tests/results/order-by-nulls-ordering.sql.out:22-96 (tests/sort_goldens.py): a sliding ROWS frame over every NULL placement,
# then the outer ORDER BY sum_col2 whose ties keep the window's output order
def test_window_under_every_null_ordering_golden(gpu, stream):
    import sort_goldens as G
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, SortExec
    t = G.t1()
    for orders, want in G.T1_WINDOW:
        w = _window(t, ["col1"], orders, [("sum", "col2", ("rows", -2, 2), 0, "sum_col2")], stream)
        got = SortExec([("sum_col2", True, True)], LocalTableScanExec(ColumnarBatch.from_arrow(w, stream))).collect(stream)
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col2", "col3", "sum_col2")])) == want, orders
