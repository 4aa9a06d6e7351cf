"""GPU parity: SortExec / TakeOrderedAndProjectExec vs the oracle -- exact output order (incl. tie order)."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def _sort(t, orders, stream):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, SortExec
    batch = ColumnarBatch.from_arrow(t, stream)
    return SortExec(orders, LocalTableScanExec(batch)).collect(stream)


def _col(kind, n, rng, null_frac):
    mask = rng.random(n) < null_frac if null_frac else None
    if kind == "int64":
        return pa.array(rng.integers(-2 ** 63, 2 ** 63 - 1, n), mask=mask)
    if kind == "int64_small":
        return pa.array(rng.integers(-20, 20, n), mask=mask)
    if kind == "int32":
        return pa.array(rng.integers(-2 ** 31, 2 ** 31, n).astype(np.int32), mask=mask)
    if kind == "int8":
        return pa.array(rng.integers(-128, 128, n).astype(np.int8), mask=mask)
    if kind == "date32":
        return pa.array(rng.integers(-1000, 20000, n).astype(np.int32), mask=mask).cast(pa.date32())
    if kind == "bool":
        return pa.array(rng.integers(0, 2, n).astype(bool), mask=mask)
    if kind in ("float64", "float32"):
        f = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
        f[rng.random(n) < 0.05] = 0.0
        f[rng.random(n) < 0.05] = -0.0
        f[rng.random(n) < 0.05] = np.nan
        f[rng.random(n) < 0.02] = np.inf
        f[rng.random(n) < 0.02] = -np.inf
        return pa.array(f.astype(np.float32 if kind == "float32" else np.float64), mask=mask)
    raise ValueError(kind)


# SortSuite.scala:36-166: every atomic type x nullable x (asc|desc) x (nulls first|last); radix path
@pytest.mark.parametrize("kind", ["int64", "int64_small", "int32", "int8", "date32", "bool", "float64", "float32"])
@pytest.mark.parametrize("null_frac", [0.0, 0.2])
@pytest.mark.parametrize("asc", [True, False])
@pytest.mark.parametrize("nulls_first", [True, False])
def test_single_column_sort_exact_order(gpu, stream, kind, null_frac, asc, nulls_first):
    n = 20000
    rng = np.random.default_rng(hash((kind, asc, nulls_first)) % 2 ** 32)
    t = pa.table({"k": _col(kind, n, rng, null_frac), "row": np.arange(n, dtype=np.int64)})
    got = _sort(t, [("k", asc, nulls_first)], stream)
    want = O.sort(t, [("k", asc, nulls_first)])
    # the payload order pins stability and the UnsafeInMemorySorter NULL-swap tie order
    assert got.column("row").to_pylist() == want.column("row").to_pylist()


@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000])
def test_sort_small_and_empty(gpu, stream, n):
    rng = np.random.default_rng(n)
    t = pa.table({"k": pa.array(rng.integers(0, 5, n), mask=rng.random(n) < 0.3, type=pa.int64()),
                  "row": np.arange(n, dtype=np.int64)})
    got = _sort(t, [("k", False, True)], stream)
    want = O.sort(t, [("k", False, True)])
    assert got.column("row").to_pylist() == want.column("row").to_pylist()


def test_null_swap_tie_order_adversarial(gpu, stream):
    """Leading NULLs, long NULL runs and a one-element queue exercise the insertRecord replay."""
    pats = [[None, None, 5, None, 5, 5, None, None, None, 5, 1, 5, None],
            [7] + [None] * 50 + [7, 7, 3],
            [None] * 10,
            [1, 2, 3] * 5 + [None] + [1, 2, 3] * 5 + [None, None] + [2] * 7]
    for vals in pats:
        n = len(vals)
        t = pa.table({"k": pa.array(vals, type=pa.int64()), "row": np.arange(n, dtype=np.int64)})
        for asc in (True, False):
            for nf in (True, False):
                got = _sort(t, [("k", asc, nf)], stream)
                want = O.sort(t, [("k", asc, nf)])
                assert got.column("row").to_pylist() == want.column("row").to_pylist(), (vals, asc, nf)


def test_multi_column_sort_matches_full_row_ordering(gpu, stream):
    n = 30000
    rng = np.random.default_rng(17)
    t = pa.table({"a": _col("int8", n, rng, 0.1), "b": _col("float64", n, rng, 0.1), "c": _col("date32", n, rng, 0.0),
                  "row": np.arange(n, dtype=np.int64)})
    for orders in ([("a", True, True), ("b", False, False)], [("a", False, False), ("c", True, True), ("b", True, False)],
                   [("b", True, True), ("a", True, False)]):
        got = _sort(t, orders, stream)
        want = O.sort(t, orders)
        assert got.column("row").to_pylist() == want.column("row").to_pylist(), orders


def test_sort_carries_payload_columns_with_nulls_and_strings(gpu, stream):
    n = 5000
    rng = np.random.default_rng(23)
    t = pa.table({"k": _col("int32", n, rng, 0.1), "p": _col("float64", n, rng, 0.2),
                  "s": pa.array([None if i % 11 == 0 else "s%d" % (i % 97) for i in range(n)], type=pa.string())})
    got = _sort(t, [("k", True, False)], stream)
    want = O.sort(t, [("k", True, False)])
    assert_tables_equal(got, want, ordered=True)


def test_take_ordered_and_project(gpu, stream):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, TakeOrderedAndProjectExec
    from spark_b200.expressions import SortOrder
    n = 100000
    rng = np.random.default_rng(31)
    t = pa.table({"revenue": rng.random(n) * 1e6, "o_orderdate": pa.array(rng.integers(8000, 9500, n).astype(np.int32)).cast(pa.date32()),
                  "l_orderkey": rng.integers(0, 10 ** 9, n)})
    batch = ColumnarBatch.from_arrow(t, stream)
    orders = [SortOrder("revenue", False), SortOrder("o_orderdate", True)]     # q3.sql: order by revenue desc, o_orderdate
    got = TakeOrderedAndProjectExec(10, orders, ["l_orderkey", "revenue", "o_orderdate"], LocalTableScanExec(batch)).collect(stream)
    want = O.take_ordered(t, [o.as_tuple() for o in orders], 10).select(["l_orderkey", "revenue", "o_orderdate"])
    assert_tables_equal(got, want, ordered=True)


def test_sort_full_size_properties(gpu, stream):
    """25M keys (SortBenchmark size): sortedness + permutation + stability, checked with numpy."""
    n = 25_000_000
    rng = np.random.default_rng(77)
    k = rng.integers(0, 1 << 20, n)                      # many ties -> stability matters
    t = pa.table({"k": k, "row": np.arange(n, dtype=np.int64)})
    got = _sort(t, [("k", True, True)], stream)
    gk = np.asarray(got.column("k")); gr = np.asarray(got.column("row"))
    assert np.all(gk[1:] >= gk[:-1])
    assert np.array_equal(gk, k[gr])
    ties = gk[1:] == gk[:-1]
    assert np.all(gr[1:][ties] > gr[:-1][ties])
    assert gr.sum() == n * (n - 1) // 2


@pytest.mark.parametrize("asc", [True, False])
@pytest.mark.parametrize("nulls_first", [True, False])
@pytest.mark.parametrize("kind", ["int64_small", "float64", "date32"])
def test_range_partition_and_global_sort(gpu, stream, kind, asc, nulls_first):
    """Global sort = RangePartitioning exchange + per-partition SortExec (SortExec.scala:54-55 requires
    OrderedDistribution; RangePartitioner.getPartition Partitioner.scala:241-260).  Partition ids and the stable regrouping
    must equal the oracle's; sorting every range and concatenating them must give a totally ordered key column."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, RangePartitioning, ShuffleExchangeExec, SortExec
    n = 60000
    rng = np.random.default_rng(hash((kind, asc, nulls_first)) % 2 ** 32)
    t = pa.table({"k": _col(kind, n, rng, 0.1), "row": np.arange(n, dtype=np.int64)})
    # bounds: 15 sample values in sort order (duplicates, a NULL and a NaN bound are legal)
    sample = t.slice(0, 15).select(["k"])
    bounds_t = O.sort(sample, [("k", asc, nulls_first)])
    bounds = bounds_t.column("k").combine_chunks()
    batch = ColumnarBatch.from_arrow(t, stream)
    bb = ColumnarBatch.from_arrow(bounds_t, stream)
    ex = ShuffleExchangeExec(RangePartitioning(("k", asc, nulls_first), bb), LocalTableScanExec(batch))
    part = ex.executeColumnar(stream)
    got = part.to_arrow(stream)
    want, offs = O.range_partition(t, ("k", asc, nulls_first), bounds)
    assert np.array_equal(ex.partition_offsets, offs)
    assert got.column("row").to_pylist() == want.column("row").to_pylist()        # stable regrouping, exact
    # per-range sort + concatenation = global order
    pieces = []
    for p in range(len(offs) - 1):
        if offs[p + 1] > offs[p]:
            sl = part.slice(int(offs[p]), int(offs[p + 1]), stream)
            pieces.append(SortExec([("k", asc, nulls_first)], LocalTableScanExec(sl)).collect(stream))
    glob = pa.concat_tables(pieces)
    ref = O.sort(t, [("k", asc, nulls_first)])
    gk, rk = glob.column("k").to_pylist(), ref.column("k").to_pylist()
    same = lambda a, b: (a is None and b is None) or (a is not None and b is not None and (a == b or (a != a and b != b)))
    assert len(gk) == len(rk) and all(same(a, b) for a, b in zip(gk, rk))
    assert sorted(glob.column("row").to_pylist()) == list(range(n))


# The key column of a single integer-typed, NULL-free order is rebuilt from the sorted 64-bit keys (csrc/sort.cu: unkey_kernel), not
# gathered: every column of the result must still equal the oracle's.
@pytest.mark.parametrize("kind", ["int64", "int64_small", "int32", "int8", "date32"])
@pytest.mark.parametrize("asc", [True, False])
@pytest.mark.parametrize("n", [1, 2, 1000, 70001])       # one row, the single-block sort (<= 2048 rows), the onesweep passes
def test_sorted_key_column_rebuilt_from_keys(gpu, stream, kind, asc, n):
    rng = np.random.default_rng(hash((kind, asc)) % 2 ** 32)
    t = pa.table({"p": rng.random(n), "k": _col(kind, n, rng, 0.0), "row": np.arange(n, dtype=np.int64)})
    assert_tables_equal(_sort(t, [("k", asc, True)], stream), O.sort(t, [("k", asc, True)]), ordered=True)
    one = pa.table({"k": _col(kind, n, rng, 0.0)})
    assert_tables_equal(_sort(one, [("k", asc, False)], stream), O.sort(one, [("k", asc, False)]), ordered=True)


# every onesweep tile geometry (sb_config_set("sort_variant")) against the oracle: ragged sizes, full-range keys (8 passes) and
# heavily duplicated keys (tie order across tiles and warps)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_onesweep_variants_exact_order(gpu, stream, variant):
    from spark_b200 import _capi as capi
    prev = capi.config_get("sort_variant")
    capi.config_set("sort_variant", variant)
    try:
        for n, hi in ((2049, 3), (4608 * 3 + 5, 2 ** 62), (300007, 700), (300007, 2 ** 62)):
            rng = np.random.default_rng(n + variant)
            t = pa.table({"k": rng.integers(-hi, hi, n), "row": np.arange(n, dtype=np.int64)})
            got = _sort(t, [("k", True, True)], stream)
            want = O.sort(t, [("k", True, True)])
            assert got.column("row").to_pylist() == want.column("row").to_pylist(), (variant, n, hi)
            assert got.column("k").to_pylist() == want.column("k").to_pylist(), (variant, n, hi)
    finally:
        capi.config_set("sort_variant", prev)


# sql-tests/results/order-by-nulls-ordering.sql.out (tests/sort_goldens.py): the reference's own answers, not the oracle's
def test_order_by_nulls_ordering_goldens(gpu, stream):
    import sort_goldens as G
    t = G.t1()
    for orders, want in G.T1_ORDER_BY:
        got = _sort(t, orders, stream)
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col2", "col3")])) == want, orders
    t = G.t2()
    for orders, want in G.T2_ORDER_BY:
        got = _sort(t, orders, stream)
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col3", "col5")])) == want, orders


def test_order_by_all_goldens(gpu, stream):
    """sql-tests/results/order-by-all.sql.out (tests/sort_goldens.py): every direction / NULL placement over two columns, and
    ORDER BY ... LIMIT 2 through TakeOrderedAndProjectExec."""
    import sort_goldens as G
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, TakeOrderedAndProjectExec
    from spark_b200.expressions import SortOrder
    t = G.t3()
    for orders, want in G.T3_ORDER_BY:
        got = _sort(t, orders, stream)
        assert list(zip(got.column("g").to_pylist(), got.column("i").to_pylist())) == want, orders
    orders, k, want = G.T3_LIMIT_2
    top = TakeOrderedAndProjectExec(k, [SortOrder(c, asc, nf) for c, asc, nf in orders], ["g", "i"],
                                    LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))).collect(stream)
    assert list(zip(top.column("g").to_pylist(), top.column("i").to_pylist())) == want
