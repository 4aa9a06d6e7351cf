"""GPU parity: SUM / AVG over decimal columns (csrc/decimal.cu) against the reference's DataFrameAggregateSuite expectations and
the oracle's exact integer restatement."""
import decimal as D

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from test_decimal_cpu import dec

pytestmark = pytest.mark.gpu


def _agg(table, keys, aggs, stream, mode="complete", batches=1):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashAggregateExec, LocalTableScanExec
    from spark_b200.expressions import Average, Count, Max, Min, Sum, col
    fn = {"sum": Sum, "avg": Average, "count": Count, "min": Min, "max": Max}
    specs = [(fn[f](col(c)), name) for f, c, name in aggs]
    if batches > 1:      # an iterator of batches through the aggregation state (Partial per batch, PartialMerge, Final)
        step = -(-table.num_rows // batches)
        parts = [ColumnarBatch.from_arrow(table.slice(i * step, step), stream) for i in range(batches)]
        partial = HashAggregateExec(keys, specs, LocalTableScanExec(parts[0]), mode="partial").execute_batches(parts, stream)
        return HashAggregateExec(keys, specs, LocalTableScanExec(partial), mode="final").collect(stream)
    scan = LocalTableScanExec(ColumnarBatch.from_arrow(table, stream))
    if mode == "complete":
        return HashAggregateExec(keys, specs, scan).collect(stream)
    return HashAggregateExec(keys, specs, HashAggregateExec(keys, specs, scan, mode="partial"), mode="final").collect(stream)


def _same(got, want, keys):
    assert got.schema.types == want.schema.types, (got.schema, want.schema)
    key = lambda r: tuple((x is None, 0 if x is None else x) for x in r[:max(1, len(keys))])
    g = sorted(zip(*[got.column(i).to_pylist() for i in range(got.num_columns)]), key=key)
    w = sorted(zip(*[want.column(i).to_pylist() for i in range(want.num_columns)]), key=key)
    assert g == w


@pytest.mark.parametrize("mode", ["complete", "partial_final"])
def test_reference_expectations(gpu, stream, mode):
    # DataFrameAggregateSuite.scala:99-112 (sum with NULLs, NULL key), :1043-1049 (avg = 1.5 per group), :2382-2386 (SPARK-36926)
    t = pa.table({"a": pa.array([1, 1, 2, 2, 3, 3, None], type=pa.int32()), "b": dec(["1", None, "1", None, "1", "2", "2"], 10, 0)})
    got = _agg(t, ["a"], [("sum", "b", "s")], stream, mode)
    rows = sorted(zip(got.column("a").to_pylist(), got.column("s").to_pylist()), key=lambda r: (r[0] is None, r[0] or 0))
    assert rows == [(1, D.Decimal(1)), (2, D.Decimal(1)), (3, D.Decimal(3)), (None, D.Decimal(2))]
    assert got.column("s").type == pa.decimal128(20, 0)
    t = pa.table({"k": pa.array([1, 1, 2, 2, 3, 3], type=pa.int32()), "b": dec(["1", "2", "1", "2", "1", "2"], 10, 2)})
    got = _agg(t, ["k"], [("avg", "b", "m")], stream, mode)
    assert sorted(got.column("m").to_pylist()) == [D.Decimal("1.500000")] * 3 and got.column("m").type == pa.decimal128(14, 6)
    t = pa.table({"d": dec(["9999999999.99"] * 10, 12, 2)})
    got = _agg(t, [], [("avg", "d", "m")], stream, mode)
    assert str(got.column("m").to_pylist()[0]) == "9999999999.990000"
    got = _agg(pa.table({"d": dec([], 10, 0)}), [], [("avg", "d", "m"), ("sum", "d", "s")], stream, mode)     # :360 empty input -> NULL
    assert got.column("m").to_pylist() == [None] and got.column("s").to_pylist() == [None]


@pytest.mark.parametrize("p,s", [(7, 2), (12, 2), (18, 4)])
@pytest.mark.parametrize("how", ["complete", "partial_final", "batches"])
def test_random_decimal_sums_and_averages(gpu, stream, p, s, how):
    """decimal(7, 2): the sum type decimal(17, 2) still fits 64 bits; (12, 2) / (18, 4): 128-bit sums.  Values use the whole precision,
    so limb sums carry across 2^32 and 2^64; NULL inputs, a group with only NULLs, negative values."""
    rng = np.random.default_rng(p * 100 + s)
    n, groups = 300_000, 1000
    lim = 10 ** p - 1
    ints = [int(x) for x in rng.integers(-lim, lim, n, dtype=np.int64)] if p <= 18 else None
    vals = [None if rng.random() < 0.05 else D.Decimal(v).scaleb(-s) for v in ints]
    k = rng.integers(0, groups, n).astype(np.int32)
    for i in np.nonzero(k == 7)[0]:
        vals[i] = None                                   # group 7: every input NULL -> NULL sum and average
    t = pa.table({"k": pa.array(k), "v": pa.array(vals, type=pa.decimal128(p, s)), "w": pa.array(rng.integers(0, 100, n), type=pa.int64())})
    aggs = [("sum", "v", "s"), ("avg", "v", "a"), ("count", "v", "c"), ("min", "v", "lo"), ("max", "v", "hi")]
    got = _agg(t, ["k"], aggs, stream, "complete" if how == "complete" else "partial_final", batches=4 if how == "batches" else 1)
    want = O.decimal_aggregate(t, ["k"], aggs)
    _same(got, want, ["k"])


def test_merge_overflow_is_null(gpu, stream):
    """Sum.scala:160-178: a merged sum beyond decimal(p + 10, s) is NULL (non-ANSI).  Hand-made Partial buffers (sum decimal(28, 0),
    isEmpty) just below 10^28 overflow when merged; an all-empty group stays NULL; a fitting one is exact."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashAggregateExec, LocalTableScanExec
    from spark_b200.expressions import Sum, col
    big = D.Decimal(9 * 10 ** 27)
    part = pa.table({"k": pa.array([1, 1, 2, 2, 3], type=pa.int32()),
                     "s#sum": pa.array([big, big, D.Decimal(5), D.Decimal(-7), D.Decimal(0)], type=pa.decimal128(28, 0)),
                     "s#isEmpty": pa.array([False, False, False, False, True])})
    got = HashAggregateExec(["k"], [(Sum(col("x")), "s")], LocalTableScanExec(ColumnarBatch.from_arrow(part, stream)), mode="final").collect(stream)
    rows = dict(zip(got.column("k").to_pylist(), got.column("s").to_pylist()))
    assert rows == {1: None, 2: D.Decimal(-2), 3: None}
