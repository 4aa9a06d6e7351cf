"""Oracle pin for decimal SUM / AVG against the reference's own expectations
(sql/core/src/test/scala/org/apache/spark/sql/DataFrameAggregateSuite.scala)."""
import decimal as D

import pyarrow as pa

from oracle import oracle as O


def dec(xs, p, s):
    return pa.array([None if x is None else D.Decimal(x) for x in xs], type=pa.decimal128(p, s))


def test_groupby_sum_of_decimals():
    # DataFrameAggregateSuite.scala:93-96 decimalData.groupBy("a").agg(sum("b")) -- DecimalData(a, b) are decimal(38, 18) in the suite; the
    # same values at decimal(10, 0) group and add identically
    t = pa.table({"a": pa.array([1, 1, 2, 2, 3, 3], type=pa.int32()), "b": dec(["1", "2", "1", "2", "1", "2"], 10, 0)})
    got = O.decimal_aggregate(t, ["a"], [("sum", "b", "s")])
    assert sorted(zip(got.column("a").to_pylist(), got.column("s").to_pylist())) == [(1, D.Decimal(3)), (2, D.Decimal(3)), (3, D.Decimal(3))]
    assert got.column("s").type == pa.decimal128(20, 0)                       # Sum.scala: decimal(p + 10, s)
    # :99-112 with NULLs: NULL inputs are skipped, a NULL key is a group
    t = pa.table({"a": pa.array([1, 1, 2, 2, 3, 3, None], type=pa.int32()), "b": dec(["1", None, "1", None, "1", "2", "2"], 10, 0)})
    got = O.decimal_aggregate(t, ["a"], [("sum", "b", "s")])
    rows = sorted(zip(got.column("a").to_pylist(), got.column("s").to_pylist()), key=lambda r: (r[0] is None, r[0] or 0))
    assert rows == [(1, D.Decimal(1)), (2, D.Decimal(1)), (3, D.Decimal(3)), (None, D.Decimal(2))]


def test_average_of_decimals():
    # :343-352 decimalData.agg(avg($"a" cast DecimalType(10, 2))) = 2 ; :1043-1049 groupBy(a).agg(avg(b cast decimal(10, 2))) = 1.5 per group
    t = pa.table({"a": dec(["1", "1", "2", "2", "3", "3"], 10, 2), "b": dec(["1", "2", "1", "2", "1", "2"], 10, 2),
                  "k": pa.array([1, 1, 2, 2, 3, 3], type=pa.int32())})
    got = O.decimal_aggregate(t, [], [("avg", "a", "m")])
    assert got.column("m").to_pylist() == [D.Decimal("2.000000")] and got.column("m").type == pa.decimal128(14, 6)   # Average.scala: (p + 4, s + 4)
    got = O.decimal_aggregate(t, ["k"], [("avg", "b", "m")])
    assert sorted(got.column("m").to_pylist()) == [D.Decimal("1.500000")] * 3
    # :2382-2386 SPARK-36926: ten times 9999999999.99 at decimal(12, 2) -> "9999999999.990000"
    t = pa.table({"d": dec(["9999999999.99"] * 10, 12, 2)})
    got = O.decimal_aggregate(t, [], [("avg", "d", "m")])
    assert str(got.column("m").to_pylist()[0]) == "9999999999.990000"
    # :360 avg over an empty input is NULL
    got = O.decimal_aggregate(pa.table({"d": dec([], 10, 0)}), [], [("avg", "d", "m"), ("sum", "d", "s")])
    assert got.column("m").to_pylist() == [None] and got.column("s").to_pylist() == [None]


def test_sum_overflow_is_null():
    # :3856-3868 SPARK-28224: 1111...1.123 + 9999...9.123 (20 integer digits, decimal(38, 18) in the suite) overflows -> NULL (non-ANSI).
    # The same digits at decimal(18, 0): two values near 10^18 whose sum is representable, and a sum beyond decimal(28, 0) is NULL
    big = "9" * 18
    t = pa.table({"a": dec([big] * 3, 18, 0)})
    got = O.decimal_aggregate(t, [], [("sum", "a", "s")])
    assert got.column("s").to_pylist() == [D.Decimal(int(big) * 3)] and got.column("s").type == pa.decimal128(28, 0)
    # rounding: HALF_UP at scale s + 4, negative values round away from zero
    t = pa.table({"a": dec(["0.01", "0.01", "0.02"], 10, 2), "b": dec(["-0.01", "-0.01", "-0.02"], 10, 2)})
    got = O.decimal_aggregate(t, [], [("avg", "a", "x"), ("avg", "b", "y")])
    assert got.column("x").to_pylist() == [D.Decimal("0.013333")] and got.column("y").to_pylist() == [D.Decimal("-0.013333")]
    t = pa.table({"a": dec(["0.00", "0.00", "0.01", "0.01", "0.01", "0.00", "0.00", "0.00"], 10, 2)})   # 0.03 / 8 = 0.00375 exactly
    assert O.decimal_aggregate(t, [], [("avg", "a", "x")]).column("x").to_pylist() == [D.Decimal("0.003750")]
