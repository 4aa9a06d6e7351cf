"""Join fixtures transcribed from the reference's suites (literal input rows and literal expected answers):
  OuterJoinSuite.scala:52-78 (left / right), :79-82 condition a = c AND b < d, answers :191-258
  ExistenceJoinSuite.scala:39-80 (left / right / conditions), answers :356-461
Used by tests/test_oracle_golden.py (the oracle against the reference's answers) and tests/test_gpu_join.py (the GPU against both)."""
import pyarrow as pa

N = None


def _t(rows, names, types):
    cols = list(zip(*rows)) if rows else [[] for _ in names]
    return pa.table({n: pa.array(list(c), type=t) for n, c, t in zip(names, cols, types)})


I, D = pa.int32(), pa.float64()
OUTER_LEFT = _t([(1, 2.0), (2, 100.0), (2, 1.0), (2, 1.0), (3, 3.0), (5, 1.0), (6, 6.0), (N, N)], ["a", "b"], [I, D])
OUTER_RIGHT = _t([(0, 0.0), (2, 3.0), (2, -1.0), (2, -1.0), (2, 3.0), (3, 2.0), (4, 1.0), (5, 3.0), (7, 7.0), (N, N)], ["c", "d"], [I, D])
COND_B_LT_D = ("lt", ("col", "b"), ("col", "d"))       # the non-equi half of `a = c AND b < d`

OUTER_CASES = {
    "left_outer": [(N, N, N, N), (1, 2.0, N, N), (2, 100.0, N, N), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0),
                   (3, 3.0, N, N), (5, 1.0, 5, 3.0), (6, 6.0, N, N)],
    "right_outer": [(N, N, N, N), (N, N, 0, 0.0), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0), (N, N, 2, -1.0), (N, N, 2, -1.0), (2, 1.0, 2, 3.0),
                    (2, 1.0, 2, 3.0), (N, N, 3, 2.0), (N, N, 4, 1.0), (5, 1.0, 5, 3.0), (N, N, 7, 7.0)],
    "full_outer": [(1, 2.0, N, N), (N, N, 2, -1.0), (N, N, 2, -1.0), (2, 100.0, N, N), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0), (2, 1.0, 2, 3.0),
                   (2, 1.0, 2, 3.0), (3, 3.0, N, N), (5, 1.0, 5, 3.0), (6, 6.0, N, N), (N, N, 0, 0.0), (N, N, 3, 2.0), (N, N, 4, 1.0),
                   (N, N, 7, 7.0), (N, N, N, N), (N, N, N, N)],
}

EXIST_LEFT = _t([(1, 2.0), (1, 2.0), (2, 1.0), (2, 1.0), (3, 3.0), (N, N), (N, 5.0), (6, N)], ["a", "b"], [I, D])
EXIST_RIGHT = _t([(2, 3.0), (2, 3.0), (3, 2.0), (4, 1.0), (N, N), (N, 5.0), (6, N)], ["c", "d"], [I, D])
EXIST_CASES = [   # (join type, residual condition, expected rows)
    ("left_semi", None, [(2, 1.0), (2, 1.0), (3, 3.0), (6, N)]),
    ("left_semi", COND_B_LT_D, [(2, 1.0), (2, 1.0)]),
    ("left_anti", None, [(1, 2.0), (1, 2.0), (N, N), (N, 5.0)]),
    ("left_anti", COND_B_LT_D, [(1, 2.0), (1, 2.0), (3, 3.0), (6, N), (N, 5.0), (N, N)]),
]


def rows_of(table):
    return list(zip(*[table.column(i).to_pylist() for i in range(table.num_columns)]))


def multiset(rows):
    return sorted(rows, key=lambda r: tuple((x is None, 0 if x is None else x) for x in r))


# ---- NOT IN (null-aware anti join, BroadcastHashJoinExec.scala:137-162): the five uncorrelated cases of
# ---- sql-tests/results/subquery/in-subquery/not-in-unit-tests-single-column.sql.out (views m(a, b), s(c, d): inputs/...single-column.sql:38-46)
NOT_IN_M = [(None, 1.0), (2, 3.0), (4, 5.0)]
NOT_IN_S = [(None, 1.0), (2, 3.0), (6, 7.0)]
# (filter on m.b or None, filter on s.d as (op, literal), expected rows of m)
NOT_IN_CASES = [
    (None, (">", 10.0), [(2, 3.0), (4, 5.0), (None, 1.0)]),   # case 1: empty subquery -> every row, NULL probe key included
    (None, ("=", 1.0), []),                                     # case 2: the subquery holds a NULL -> no row
    (1.0, ("=", 3.0), []),                                      # case 3: the probe key is NULL -> not returned
    (3.0, ("=", 3.0), []),                                      # case 4: the probe key is in the subquery -> not returned
    (3.0, ("=", 7.0), [(2, 3.0)]),                              # case 5: the probe key is not in the subquery -> returned
]
