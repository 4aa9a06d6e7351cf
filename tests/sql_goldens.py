"""Aggregate answers transcribed from the reference's SQL golden files (sql/core/src/test/resources/sql-tests):
  inputs/group-by.sql:7-9 testData, :94-99 test_agg; results/group-by.sql.out (query -> rows), results/having.sql.out.
Each case = (name, table, group keys (expressions over the table, None = global), aggregates, post filter, expected rows as the
golden file prints them, file:line of the expected output).  The plan for each query is written by hand (there is no SQL parser
on this path): the physical operators are what the answers pin."""
import numpy as np
import pyarrow as pa

N = None
TEST_DATA = pa.table({"a": pa.array([1, 1, 2, 2, 3, 3, N, 3, N], type=pa.int32()), "b": pa.array([1, 2, 1, 2, 1, 2, 1, N, N], type=pa.int32())})
TEST_AGG = pa.table({"k": pa.array([1, 1, 2, 3, 3, 4, 4, 5, 5, 5], type=pa.int32()),
                     "v": pa.array([True, False, True, False, N, N, N, N, True, False], type=pa.bool_())})
HAV = pa.table({"k": pa.array([1, 2, 3, 1], type=pa.int32()),        # 'one' -> 1, 'two' -> 2, 'three' -> 3 (string keys as codes)
                "v": pa.array([1, 2, 3, 5], type=pa.int32())})

A, B = ("col", "a"), ("col", "b")
CASES = [
    # group-by.sql.out:40-44   SELECT COUNT(a), COUNT(b) FROM testData
    ("count_global", TEST_DATA, None, [], [("count", "a", "ca"), ("count", "b", "cb")], None, [(7, 7)], "group-by.sql.out:44"),
    # :48-55   SELECT a, COUNT(b) FROM testData GROUP BY a
    ("count_by_a", TEST_DATA, ["a"], [], [("count", "b", "cb")], None, [(1, 2), (2, 2), (3, 2), (N, 1)], "group-by.sql.out:52-55"),
    # :75-82   SELECT COUNT(a), COUNT(b) FROM testData GROUP BY a
    ("counts_by_a", TEST_DATA, ["a"], [], [("count", "a", "ca"), ("count", "b", "cb")], "drop_keys", [(0, 1), (2, 2), (2, 2), (3, 2)],
     "group-by.sql.out:79-82"),
    # :118-126 SELECT a + b, COUNT(b) FROM testData GROUP BY a + b
    ("count_by_a_plus_b", TEST_DATA, ["g"], [("g", ("add", A, B))], [("count", "b", "cb")], None, [(2, 1), (3, 2), (4, 2), (5, 1), (N, 1)],
     "group-by.sql.out:122-126"),
    # :146-153 SELECT a + 1 + 1, COUNT(b) FROM testData GROUP BY a + 1
    ("count_by_a_plus_1", TEST_DATA, ["g"], [("g", ("add", ("add", A, ("lit", 1, np.int32)), ("lit", 1, np.int32)))], [("count", "b", "cb")], None,
     [(3, 2), (4, 2), (5, 2), (N, 1)], "group-by.sql.out:150-153"),
    # :168-173 MIN(a), MAX(a), AVG(a), SUM(a), COUNT(a) FROM testData  (skewness / variance columns are not on this path)
    ("min_max_avg_sum_count", TEST_DATA, None, [], [("min", "a", "mn"), ("max", "a", "mx"), ("avg", "a", "av"), ("sum", "a", "s"), ("count", "a", "c")],
     None, [(1, 3, 2.142857142857143, 15, 7)], "group-by.sql.out:173"),
    # :218-222 SELECT a, COUNT(1) FROM testData WHERE false GROUP BY a   -> no rows
    ("empty_input_grouped", TEST_DATA.slice(0, 0), ["a"], [], [("count_star", None, "c")], None, [], "group-by.sql.out:222"),
    # :226-230 SELECT COUNT(1) FROM testData WHERE false   -> one row: 0
    ("empty_input_global", TEST_DATA.slice(0, 0), None, [], [("count_star", None, "c")], None, [(0,)], "group-by.sql.out:230"),
    # :692-696 SELECT count(*) FROM test_agg HAVING count(*) > 1L
    ("count_star_having", TEST_AGG, None, [], [("count_star", None, "c")], ("gt", ("col", "c"), ("lit", 1, np.int64)), [(10,)], "group-by.sql.out:696"),
    # :700-706 SELECT k, max(v) FROM test_agg GROUP BY k HAVING max(v) = true
    ("max_bool_having", TEST_AGG, ["k"], [], [("max", "v", "m")], ("eq", ("col", "m"), ("lit", True, np.bool_)), [(1, True), (2, True), (5, True)],
     "group-by.sql.out:704-706"),
    # having.sql.out:16-21 SELECT k, sum(v) FROM hav GROUP BY k HAVING sum(v) > 2   ('one' 6, 'three' 3)
    ("sum_having", HAV, ["k"], [], [("sum", "v", "s")], ("gt", ("col", "s"), ("lit", 2, np.int64)), [(1, 6), (3, 3)], "having.sql.out:20-21"),
    # having.sql.out:57-61 SELECT MIN(t.v) FROM (SELECT * FROM hav WHERE v > 0) t HAVING(COUNT(1) > 0)
    ("min_having_count", HAV, None, [], [("min", "v", "m"), ("count_star", None, "c")], ("gt", ("col", "c"), ("lit", 0, np.int64)), [(1, 4)],
     "having.sql.out:61"),
]


def norm(rows):
    return sorted(rows, key=lambda r: tuple((x is None, 0 if x is None else x) for x in r))
