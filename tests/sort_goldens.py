"""ORDER BY answers held by the reference: sql/core/src/test/resources/sql-tests/results/order-by-nulls-ordering.sql.out (SPARK-10747),
transcribed query by query.  The rows are the INSERT statements of the file (:10-11 and :172-181), the expected outputs the
`-- !query output` blocks.  col5 of the second table is decimal(20,1) in the reference; the sort path takes decimals up to 18 digits,
so it is held here as a double with the same values (the ordering of 10.0, 0.0, 15.1, 1.0, NULL does not depend on the type)."""
import pyarrow as pa

N = None

# spark_10747(col1 int, col2 int, col3 int)                                                        -- order-by-nulls-ordering.sql.out:10-11
T1_ROWS = [(6, 12, 10), (6, 11, 4), (6, 9, 10), (6, 15, 8), (6, 15, 8), (6, 7, 4), (6, 7, 8), (6, 13, N), (6, 10, N)]


def t1():
    c1, c2, c3 = zip(*T1_ROWS)
    return pa.table({"col1": pa.array(c1, pa.int32()), "col2": pa.array(c2, pa.int32()), "col3": pa.array(c3, pa.int32())})


# (orders as (column, ascending, nulls_first), expected rows)                                       -- :98-160
T1_ORDER_BY = [
    ([("col3", True, True), ("col2", True, True)],                      # ORDER BY COL3 ASC NULLS FIRST, COL2
     [(6, 10, N), (6, 13, N), (6, 7, 4), (6, 11, 4), (6, 7, 8), (6, 15, 8), (6, 15, 8), (6, 9, 10), (6, 12, 10)]),
    ([("col3", True, False), ("col2", True, True)],                     # ORDER BY COL3 NULLS LAST, COL2
     [(6, 7, 4), (6, 11, 4), (6, 7, 8), (6, 15, 8), (6, 15, 8), (6, 9, 10), (6, 12, 10), (6, 10, N), (6, 13, N)]),
    ([("col3", False, True), ("col2", True, True)],                     # ORDER BY COL3 DESC NULLS FIRST, COL2
     [(6, 10, N), (6, 13, N), (6, 9, 10), (6, 12, 10), (6, 7, 8), (6, 15, 8), (6, 15, 8), (6, 7, 4), (6, 11, 4)]),
    ([("col3", False, False), ("col2", True, True)],                    # ORDER BY COL3 DESC NULLS LAST, COL2
     [(6, 9, 10), (6, 12, 10), (6, 7, 8), (6, 15, 8), (6, 15, 8), (6, 7, 4), (6, 11, 4), (6, 10, N), (6, 13, N)]),
]

# sum(col2) over (partition by col1 order by col3 <dir> nulls <place>, col2 rows between 2 preceding and 2 following) ... order by sum_col2
# (window order, expected (col1, col2, col3, sum_col2) rows)                                         -- :22-96
T1_WINDOW = [
    ([("col3", False, False), ("col2", True, True)],
     [(6, 9, 10, 28), (6, 13, N, 34), (6, 10, N, 41), (6, 12, 10, 43), (6, 15, 8, 55), (6, 15, 8, 56), (6, 11, 4, 56), (6, 7, 8, 58), (6, 7, 4, 58)]),
    ([("col3", False, True), ("col2", True, True)],
     [(6, 10, N, 32), (6, 11, 4, 33), (6, 13, N, 44), (6, 7, 4, 48), (6, 9, 10, 51), (6, 15, 8, 55), (6, 12, 10, 56), (6, 15, 8, 56), (6, 7, 8, 58)]),
    ([("col3", True, False), ("col2", True, True)],
     [(6, 7, 4, 25), (6, 13, N, 35), (6, 11, 4, 40), (6, 10, N, 44), (6, 7, 8, 55), (6, 15, 8, 57), (6, 15, 8, 58), (6, 12, 10, 59), (6, 9, 10, 61)]),
    ([("col3", True, True), ("col2", True, True)],
     [(6, 10, N, 30), (6, 12, 10, 36), (6, 13, N, 41), (6, 7, 4, 48), (6, 9, 10, 51), (6, 11, 4, 53), (6, 7, 8, 55), (6, 15, 8, 57), (6, 15, 8, 58)]),
]

# spark_10747_mix(col1 string, col2 int, col3 double, col4 decimal(10,2), col5 decimal(20,1))      -- :172-181
T2_ROWS = [("b", 2, 1.0, "1.00", 10.0), ("d", 3, 2.0, "3.00", 0.0), ("c", 3, 2.0, "2.00", 15.1), ("d", 3, 0.0, "3.00", 1.0),
           (N, 3, 0.0, "3.00", 1.0), ("d", 3, N, "4.00", 1.0), ("a", 1, 1.0, "1.00", N), ("c", 3, 2.0, "2.00", N)]


def t2():
    c1, c2, c3, c4, c5 = zip(*T2_ROWS)
    return pa.table({"col1": pa.array(c1, pa.string()), "col2": pa.array(c2, pa.int32()), "col3": pa.array(c3, pa.float64()),
                     "col4": pa.array([float(x) for x in c4], pa.float64()), "col5": pa.array(c5, pa.float64())})


# (orders, expected col1 / col3 / col5 of every output row)                                         -- :188-228
T2_ORDER_BY = [
    ([("col1", True, False), ("col5", True, False)],                    # order by col1 nulls last, col5 nulls last
     [("a", 1.0, N), ("b", 1.0, 10.0), ("c", 2.0, 15.1), ("c", 2.0, N), ("d", 2.0, 0.0), ("d", 0.0, 1.0), ("d", N, 1.0), (N, 0.0, 1.0)]),
    ([("col1", False, True), ("col5", False, True)],                    # order by col1 desc nulls first, col5 desc nulls first
     [(N, 0.0, 1.0), ("d", 0.0, 1.0), ("d", N, 1.0), ("d", 2.0, 0.0), ("c", 2.0, N), ("c", 2.0, 15.1), ("b", 1.0, 10.0), ("a", 1.0, N)]),
    ([("col5", False, True), ("col3", False, False)],                   # order by col5 desc nulls first, col3 desc nulls last
     [("c", 2.0, N), ("a", 1.0, N), ("c", 2.0, 15.1), ("b", 1.0, 10.0), ("d", 0.0, 1.0), (N, 0.0, 1.0), ("d", N, 1.0), ("d", 2.0, 0.0)]),
]


# sql-tests/results/order-by-all.sql.out: data(g, i) = (0, 1), (0, 2), (1, 3), (1, NULL); ORDER BY ALL = both columns with the same
# direction / NULL placement (defaults: ASC NULLS FIRST, DESC NULLS LAST -- SortOrder.scala NullOrdering defaults)        -- inputs/order-by-all.sql:1-6
T3_ROWS = [(0, 1), (0, 2), (1, 3), (1, N)]


def t3():
    g, i = zip(*T3_ROWS)
    return pa.table({"g": pa.array(g, pa.int32()), "i": pa.array(i, pa.int32())})


def _all(asc, nulls_first):
    return [("g", asc, nulls_first), ("i", asc, nulls_first)]


T3_ORDER_BY = [                                                                                    # order-by-all.sql.out:30-118
    (_all(True, True), [(0, 1), (0, 2), (1, N), (1, 3)]),        # order by all / all asc / all nulls first / all asc nulls first
    (_all(False, False), [(1, 3), (1, N), (0, 2), (0, 1)]),      # order by all desc / all desc nulls last
    (_all(True, False), [(0, 1), (0, 2), (1, 3), (1, N)]),       # order by all nulls last / all asc nulls last
    (_all(False, True), [(1, N), (1, 3), (0, 2), (0, 1)]),       # order by all desc nulls first
]
T3_LIMIT_2 = (_all(True, True), 2, [(0, 1), (0, 2)])             # order by all limit 2 (TakeOrderedAndProjectExec)      -- :163-168
