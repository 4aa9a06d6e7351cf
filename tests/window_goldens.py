"""Answers transcribed from the reference's golden files for the f4 operators:
  sql/core/src/test/resources/sql-tests/results/window.sql.out          (WindowExec)
  sql/core/src/test/resources/sql-tests/results/group-analytics.sql.out (ExpandExec below an aggregate: CUBE / ROLLUP)
Each case: the input view as the golden file creates it, the physical operator arguments the query plans to, and the rows of
the `-- !query output` block (the queries end in ORDER BY cate, val -- ties keep the golden file's order, and the comparison is
on multisets)."""
import pyarrow as pa

N = None

# window.sql.out:3-13  testData(val, val_long, val_double, val_date, val_timestamp, cate) -- the columns the cases below use
WINDOW_TEST_DATA = pa.table({
    "val": pa.array([N, 1, 1, 2, 1, 2, 3, N, 3], type=pa.int32()),
    "val_long": pa.array([1, 1, 2, 2147483650, N, 3, 2147483650, N, 1], type=pa.int64()),
    "val_double": pa.array([1.0, 1.0, 2.5, 100.001, 1.0, 3.3, 100.001, N, 1.0], type=pa.float64()),
    "cate": pa.array(["a", "a", "a", "a", "b", "b", "b", N, N], type=pa.string()),
})

# (golden line, partitionSpec, orderSpec [(col, asc, nulls_first)], [(func, col, frame, param, name)], expected rows (val, cate, results...))
# the first two output columns are (val, cate) unless a case says otherwise in WINDOW_CASE_COLUMNS
WINDOW_CASE_COLUMNS = {159: ["val_long", "cate"], 176: ["val_double", "cate"]}
WINDOW_CASES = [
    # :65  count(val) OVER(PARTITION BY cate ORDER BY val ROWS CURRENT ROW)
    (65, ["cate"], [("val", True, True)], [("count", "val", ("rows", 0, 0), 0, "c")],
     [(N, N, 0), (3, N, 1), (N, "a", 0), (1, "a", 1), (1, "a", 1), (2, "a", 1), (1, "b", 1), (2, "b", 1), (3, "b", 1)]),
    # :82  sum(val) OVER(PARTITION BY cate ORDER BY val ROWS BETWEEN UNBOUNDED PRECEDING AND 1 FOLLOWING)
    (82, ["cate"], [("val", True, True)], [("sum", "val", ("rows", None, 1), 0, "s")],
     [(N, N, 3), (3, N, 3), (N, "a", 1), (1, "a", 2), (1, "a", 4), (2, "a", 4), (1, "b", 3), (2, "b", 6), (3, "b", 6)]),
    # :520 WINDOW w AS (PARTITION BY cate ORDER BY val): the default frame RANGE UNBOUNDED PRECEDING .. CURRENT ROW
    #      columns: max, min, count, sum, avg, first_value, last_value, rank, dense_rank, cume_dist, percent_rank, ntile(2), row_number
    (520, ["cate"], [("val", True, True)],
     [("max", "val", None, 0, "max"), ("min", "val", None, 0, "min"), ("count", "val", None, 0, "count"), ("sum", "val", None, 0, "sum"),
      ("avg", "val", None, 0, "avg"), ("first_value", "val", None, 0, "first_value"), ("last_value", "val", None, 0, "last_value"),
      ("rank", None, None, 0, "rank"), ("dense_rank", None, None, 0, "dense_rank"), ("cume_dist", None, None, 0, "cume_dist"),
      ("percent_rank", None, None, 0, "percent_rank"), ("ntile", None, None, 2, "ntile"), ("row_number", None, None, 0, "row_number")],
     [(N, N, N, N, 0, N, N, N, N, 1, 1, 0.5, 0.0, 1, 1),
      (3, N, 3, 3, 1, 3, 3.0, N, 3, 2, 2, 1.0, 1.0, 2, 2),
      (N, "a", N, N, 0, N, N, N, N, 1, 1, 0.25, 0.0, 1, 1),
      (1, "a", 1, 1, 2, 2, 1.0, N, 1, 2, 2, 0.75, 0.3333333333333333, 1, 2),
      (1, "a", 1, 1, 2, 2, 1.0, N, 1, 2, 2, 0.75, 0.3333333333333333, 2, 3),
      (2, "a", 2, 1, 3, 4, 1.3333333333333333, N, 2, 4, 3, 1.0, 1.0, 2, 4),
      (1, "b", 1, 1, 1, 1, 1.0, 1, 1, 1, 1, 0.3333333333333333, 0.0, 1, 1),
      (2, "b", 2, 1, 2, 3, 1.5, 1, 2, 2, 2, 0.6666666666666666, 0.5, 1, 2),
      (3, "b", 3, 1, 3, 6, 2.0, 1, 3, 3, 3, 1.0, 1.0, 2, 3)]),
    # :125 count(val) OVER(PARTITION BY cate ORDER BY val RANGE 1 PRECEDING): value offsets over the ORDER BY key
    (125, ["cate"], [("val", True, True)], [("count", "val", ("range", -1, 0), 0, "c")],
     [(N, N, 0), (3, N, 1), (N, "a", 0), (1, "a", 2), (1, "a", 2), (2, "a", 3), (1, "b", 1), (2, "b", 2), (3, "b", 2)]),
    # :142 sum(val) OVER(PARTITION BY cate ORDER BY val RANGE BETWEEN CURRENT ROW AND 1 FOLLOWING)
    (142, ["cate"], [("val", True, True)], [("sum", "val", ("range", 0, 1), 0, "s")],
     [(N, N, N), (3, N, 3), (N, "a", N), (1, "a", 4), (1, "a", 4), (2, "a", 2), (1, "b", 3), (2, "b", 5), (3, "b", 3)]),
    # :362 the same with ORDER BY val DESC: FOLLOWING means smaller values
    (362, ["cate"], [("val", False, False)], [("sum", "val", ("range", 0, 1), 0, "s")],
     [(N, N, N), (3, N, 3), (N, "a", N), (1, "a", 2), (1, "a", 2), (2, "a", 4), (1, "b", 1), (2, "b", 3), (3, "b", 5)]),
    # :159 sum(val_long) OVER(PARTITION BY cate ORDER BY val_long RANGE BETWEEN CURRENT ROW AND 2147483648 FOLLOWING): a long offset
    (159, ["cate"], [("val_long", True, True)], [("sum", "val_long", ("range", 0, 2147483648), 0, "s")],
     [(N, N, N), (1, N, 1), (1, "a", 4), (1, "a", 4), (2, "a", 2147483652), (2147483650, "a", 2147483650), (N, "b", N), (3, "b", 2147483653),
      (2147483650, "b", 2147483650)]),
    # :176 sum(val_double) OVER(PARTITION BY cate ORDER BY val_double RANGE BETWEEN CURRENT ROW AND 2.5 FOLLOWING): a double offset
    (176, ["cate"], [("val_double", True, True)], [("sum", "val_double", ("range", 0, 2.5), 0, "s")],
     [(N, N, N), (1.0, N, 1.0), (1.0, "a", 4.5), (1.0, "a", 4.5), (2.5, "a", 2.5), (100.001, "a", 100.001), (1.0, "b", 4.3), (3.3, "b", 3.3),
      (100.001, "b", 100.001)]),
    # :621 sum(val) OVER(), avg(val) OVER(): one partition, the whole-partition frame
    (621, [], [], [("sum", "val", None, 0, "s"), ("avg", "val", None, 0, "a")],
     [(N, N, 13, 1.8571428571428572), (3, N, 13, 1.8571428571428572), (N, "a", 13, 1.8571428571428572), (1, "a", 13, 1.8571428571428572),
      (1, "a", 13, 1.8571428571428572), (2, "a", 13, 1.8571428571428572), (1, "b", 13, 1.8571428571428572), (2, "b", 13, 1.8571428571428572),
      (3, "b", 13, 1.8571428571428572)]),
]

# group-analytics.sql.out:3-5  testData(a, b)
EXPAND_TEST_DATA = pa.table({"a": pa.array([1, 1, 2, 2, 3, 3], type=pa.int32()), "b": pa.array([1, 2, 1, 2, 1, 2], type=pa.int32())})

# (golden line, grouping sets as bit masks over (g0, g1): bit set = column nulled (spark_grouping_id), g0 expr, g1 expr, agg input expr, expected rows)
#   the plan: Expand [ (agg input, g0 or NULL, g1 or NULL, gid) per grouping set ] -> HashAggregate(keys g0, g1, gid; sum(agg input))
EXPAND_CASES = [
    (13, "cube", "a+b", "b", "a-b",
     [(2, 1, 0), (2, N, 0), (3, 1, 1), (3, 2, -1), (3, N, 0), (4, 1, 2), (4, 2, 0), (4, N, 2), (5, 2, 1), (5, N, 1), (N, 1, 3), (N, 2, 0), (N, N, 3)]),
    (33, "cube", "a", "b", "b",
     [(1, 1, 1), (1, 2, 2), (1, N, 3), (2, 1, 1), (2, 2, 2), (2, N, 3), (3, 1, 1), (3, 2, 2), (3, N, 3), (N, 1, 3), (N, 2, 6), (N, N, 9)]),
    (52, "rollup", "a+b", "b", "a-b",
     [(2, 1, 0), (2, N, 0), (3, 1, 1), (3, 2, -1), (3, N, 0), (4, 1, 2), (4, 2, 0), (4, N, 2), (5, 2, 1), (5, N, 1), (N, N, 3)]),
    (70, "rollup", "a", "b", "b",
     [(1, 1, 1), (1, 2, 2), (1, N, 3), (2, 1, 1), (2, 2, 2), (2, N, 3), (3, 1, 1), (3, 2, 2), (3, N, 3), (N, N, 9)]),
]
