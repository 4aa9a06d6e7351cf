"""Pins the CPU oracle against the reference's own known-answer vectors.  CPU only.

Every expected value below is quoted from the reference tree (file:line given per test), not computed here.
"""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O


# --- Murmur3_x86_32Suite.java:38-53 (seed 0) ------------------------------------------------------
def test_murmur3_hash_int_known_answers():
    L = O.lib()
    expect = {0: 593689054, -42: -189366624, 42: -1134849565, -2 ** 31: -1718298732, 2 ** 31 - 1: -1653689534}
    for v, h in expect.items():
        assert L.so_murmur3_int(v, 0) == h


def test_murmur3_hash_long_known_answers():
    L = O.lib()
    expect = {0: 1669671676, -42: -846261623, 42: 1871679806, -2 ** 63: 1366273829, 2 ** 63 - 1: -2106506049}
    for v, h in expect.items():
        assert L.so_murmur3_long(v, 0) == h


# --- hash.scala:844-845: hash('Spark', array(123), 2) = -1321691492 (seed 42, chained, array elements chained)
def test_hash_expression_example():
    L = O.lib()
    h = L.so_murmur3_bytes(b"Spark", 5, 42)
    h = L.so_murmur3_int(123, h)
    h = L.so_murmur3_int(2, h)
    assert h == -1321691492


# --- python/pyspark/sql/functions/builtin.py:15372-15390 doctest ---------------------------------------
def test_pyspark_hash_doctest():
    t = pa.table({"c1": ["ABC"], "c2": ["DEF"]})
    assert O.hash_rows(t, ["c1"])[0] == -757602832
    assert O.hash_rows(t, ["c1", "c2"])[0] == 599895104


def test_hash_null_keeps_seed_and_negative_zero():
    # hash.scala:714 (null -> seed), :718-723 (-0.0 hashes like 0.0)
    t = pa.table({"a": pa.array([None, 1], type=pa.int32()), "d": pa.array([-0.0, 0.0])})
    h = O.hash_rows(t, ["a"])
    assert h[0] == 42
    hd = O.hash_rows(t, ["d"])
    assert hd[0] == hd[1] == O.lib().so_murmur3_long(0, 42)


def test_string_hash_legacy_tail_is_signed_per_byte():
    # Murmur3_x86_32.java:66-78: tail bytes are sign-extended and mixed one by one (differs from hashUnsafeBytes2)
    L = O.lib()
    b = bytes([0xE4, 0xBD, 0xA0])        # 3 bytes, all >= 0x80
    h1 = 42
    for byte in b:
        sb = byte - 256
        k1 = (sb * 0xcc9e2d51) & 0xFFFFFFFF
        k1 = ((k1 << 15) | (k1 >> 17)) & 0xFFFFFFFF
        k1 = (k1 * 0x1b873593) & 0xFFFFFFFF
        h1 ^= k1
        h1 = ((h1 << 13) | (h1 >> 19)) & 0xFFFFFFFF
        h1 = (h1 * 5 + 0xe6546b64) & 0xFFFFFFFF
    h1 ^= 3
    h1 ^= h1 >> 16
    h1 = (h1 * 0x85ebca6b) & 0xFFFFFFFF
    h1 ^= h1 >> 13
    h1 = (h1 * 0xc2b2ae35) & 0xFFFFFFFF
    h1 ^= h1 >> 16
    want = h1 - (1 << 32) if h1 >= 1 << 31 else h1
    assert L.so_murmur3_bytes(b, 3, 42) == want


def test_pmod_is_non_negative():
    # MathUtils.scala:96-99
    t = pa.table({"k": np.arange(-500, 500, dtype=np.int64)})
    for n in (1, 2, 7, 200, 2048):
        pid = O.partition_ids(t, ["k"], n)
        assert pid.min() >= 0 and pid.max() < n
        h = O.hash_rows(t, ["k"]).astype(np.int64)
        assert np.array_equal(pid, np.mod(h, n))


# --- RadixSortSuite.scala:45-199: radix result must equal a reference comparison sort for every sort type ----
RADIX_CONFIGS = [  # (name, start_byte, end_byte, desc, signed) as in RadixSortSuite.scala:45-73
    ("unsigned binary data asc", 0, 7, False, False), ("unsigned binary data desc", 0, 7, True, False),
    ("twos complement asc", 0, 7, False, True), ("twos complement desc", 0, 7, True, True),
    ("positive twos complement asc", 0, 7, False, True), ("positive twos complement desc", 0, 7, True, True),
    ("unsigned partial asc", 2, 4, False, False), ("unsigned partial desc", 2, 4, True, False),
]


@pytest.mark.parametrize("cfg", RADIX_CONFIGS, ids=[c[0] for c in RADIX_CONFIGS])
@pytest.mark.parametrize("n", [0, 1, 2, 1000, 65537])
def test_radix_sort_matches_reference_sort(cfg, n):
    name, sb, eb, desc, sgn = cfg
    rng = np.random.default_rng(123)
    a = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    if "positive" in name:
        a = np.abs(a // 2)
    if "partial" in name:
        a = a & np.int64(0x000000FFFFFF0000)
    got = a.copy()
    O.lib().so_radix_sort(got.ctypes.data, n, sb, eb, int(desc), int(sgn))
    key = a if sgn else a.view(np.uint64)
    want = np.sort(key)
    if desc:
        want = want[::-1]
    assert np.array_equal(got.view(key.dtype), want)


def test_radix_sort_random_bitmask_fuzz():
    # RadixSortSuite.scala:160-199: random bit masks exercise the skip-constant-bytes pre-pass
    rng = np.random.default_rng(123)
    for _ in range(10):
        mask = rng.integers(-2 ** 63, 2 ** 63 - 1, dtype=np.int64)
        a = rng.integers(-2 ** 63, 2 ** 63 - 1, 5000, dtype=np.int64) & mask
        got = a.copy()
        O.lib().so_radix_sort(got.ctypes.data, len(a), 0, 7, 0, 1)
        assert np.array_equal(got, np.sort(a))


def test_key_prefix_radix_sort_is_stable():
    # RadixSortSuite key-prefix variant compares whole (pointer, prefix) arrays => stability is pinned
    rng = np.random.default_rng(123)
    n = 20000
    prefix = rng.integers(0, 50, n, dtype=np.int64)
    ptr = np.arange(n, dtype=np.int64)
    p2, k2 = ptr.copy(), prefix.copy()
    O.lib().so_radix_sort_key_prefix(p2.ctypes.data, k2.ctypes.data, n, 0, 7, 0, 1)
    order = np.argsort(prefix, kind="stable")
    assert np.array_equal(p2, order) and np.array_equal(k2, prefix[order])
    p3, k3 = ptr.copy(), prefix.copy()
    O.lib().so_radix_sort_key_prefix(p3.ctypes.data, k3.ctypes.data, n, 0, 7, 1, 1)
    # descending walks buckets in reverse but keeps insertion order inside a bucket (RadixSort.java:154-159)
    order_desc = np.argsort(-prefix, kind="stable")
    assert np.array_equal(p3, order_desc)


# --- PrefixComparatorsSuite: double prefixes order like doubles, NaN largest, -0.0 == 0.0 ------------------
def test_double_prefix_ordering():
    L = O.lib()
    vals = [float("-inf"), -1e300, -1.5, -0.0, 0.0, 5e-324, 1.5, 1e300, float("inf"), float("nan")]
    pre = [np.uint64(L.so_double_prefix(v) & 0xFFFFFFFFFFFFFFFF) for v in vals]
    assert pre[3] == pre[4]                      # -0.0 and 0.0 share a prefix (PrefixComparators.java:69)
    dedup = pre[:3] + pre[4:]
    assert all(dedup[i] < dedup[i + 1] for i in range(len(dedup) - 1))   # unsigned compare; NaN is the largest


# --- SortSuite.scala:36-166 rules: null placement independent of direction on the radix path ------------------
@pytest.mark.parametrize("asc", [True, False])
@pytest.mark.parametrize("nulls_first", [True, False])
def test_sort_single_column_null_placement(asc, nulls_first):
    rng = np.random.default_rng(5)
    vals = rng.integers(-100, 100, 500)
    t = pa.table({"k": pa.array(vals, mask=rng.random(500) < 0.2), "row": np.arange(500)})
    out = O.sort(t, [("k", asc, nulls_first)])
    k = out.column("k").to_pylist()
    nn = [x for x in k if x is not None]
    nnull = len(k) - len(nn)
    assert (k[:nnull] if nulls_first else k[len(nn):]) == [None] * nnull
    assert nn == sorted(nn, reverse=not asc)


def test_sort_multi_column_is_stable_full_ordering():
    rng = np.random.default_rng(6)
    t = pa.table({"a": rng.integers(0, 5, 300), "b": pa.array(rng.random(300), mask=rng.random(300) < 0.1),
                  "row": np.arange(300)})
    out = O.sort(t, [("a", False, False), ("b", True, True)])
    rows = list(zip(out.column("a").to_pylist(), out.column("b").to_pylist(), out.column("row").to_pylist()))
    want = sorted(zip(t.column("a").to_pylist(), t.column("b").to_pylist(), t.column("row").to_pylist()),
                  key=lambda r: (-r[0], r[1] is not None, r[1] if r[1] is not None else 0.0, r[2]))
    assert rows == want


# --- InnerJoinSuite.scala:40-75 fixtures ------------------------------------------------------------------
def _join_fixtures():
    # myUpperCaseData / myLowerCaseData (InnerJoinSuite.scala:40-64) incl. NULL keys on both sides
    upper = pa.table({"N": pa.array([1, 2, 3, 4, 5, 6, None], type=pa.int32()),
                      "L": pa.array([0, 1, 2, 3, 4, 5, 6], type=pa.int32())})    # letters A..G as codes
    lower = pa.table({"n": pa.array([1, 2, 3, 4, None], type=pa.int32()),
                      "l": pa.array([10, 11, 12, 13, 14], type=pa.int32())})
    return upper, lower


def test_inner_join_fixture_null_keys_never_match():
    upper, lower = _join_fixtures()
    out = O.hash_join(upper, lower, ["N"], ["n"], "inner")
    rows = sorted(zip(*[out.column(i).to_pylist() for i in range(4)]))
    assert rows == [(1, 0, 1, 10), (2, 1, 2, 11), (3, 2, 3, 12), (4, 3, 4, 13)]    # InnerJoinSuite.scala:164-171


def test_outer_semi_anti_join_fixture():
    upper, lower = _join_fixtures()
    lo = O.hash_join(upper, lower, ["N"], ["n"], "left_outer")
    assert lo.num_rows == 7
    unmatched = [r for r in zip(*[lo.column(i).to_pylist() for i in range(4)]) if r[2] is None]
    assert sorted(unmatched, key=lambda r: (r[0] is None, r[0])) == [(5, 4, None, None), (6, 5, None, None), (None, 6, None, None)]
    semi = O.hash_join(upper, lower, ["N"], ["n"], "left_semi")
    assert sorted(semi.column("N").to_pylist()) == [1, 2, 3, 4]
    anti = O.hash_join(upper, lower, ["N"], ["n"], "left_anti")
    assert sorted(anti.column("L").to_pylist()) == [4, 5, 6]      # NULL-key row is kept by anti join


def test_join_duplicate_build_keys_emit_all_matches():
    # InnerJoinSuite.scala:228-299 "inner join, multiple matches": (1,1),(1,2) x (1,1),(1,2) on a -> 4 rows
    left = pa.table({"a": pa.array([1, 1, 2, 2, 3], type=pa.int32()), "b": pa.array([1, 2, 1, 2, 2], type=pa.int32())})
    right = left.rename_columns(["a2", "b2"])
    out = O.hash_join(left.filter(pa.array([True, True, False, False, False])),
                      right.filter(pa.array([True, True, False, False, False])), ["a"], ["a2"], "inner")
    rows = sorted(zip(*[out.column(i).to_pylist() for i in range(4)]))
    assert rows == [(1, 1, 1, 1), (1, 1, 1, 2), (1, 2, 1, 1), (1, 2, 1, 2)]


# --- aggregate algebra (Sum.scala:113-178, Average.scala:80-135, Count.scala:94-105) ----------------------------
def test_aggregate_null_semantics():
    t = pa.table({"k": pa.array([1, 1, 2, 2, None, None], type=pa.int32()),
                  "v": pa.array([10, None, None, None, 5, 7], type=pa.int64()),
                  "d": pa.array([1.5, 2.5, None, None, None, 4.0])})
    out = O.hash_aggregate(t, ["k"], [("sum", "v", "s"), ("count", "v", "c"), ("count_star", None, "n"),
                                       ("avg", "d", "a"), ("min", "v", "mn"), ("max", "d", "mx")])
    rows = {r[0]: r[1:] for r in zip(*[out.column(i).to_pylist() for i in range(out.num_columns)])}
    assert rows[1] == (10, 1, 2, 2.0, 10, 2.5)
    assert rows[2] == (None, 0, 2, None, None, None)       # all-NULL group: sum/avg/min/max NULL, count 0
    assert rows[None] == (12, 2, 2, 4.0, 5, 4.0)           # NULL is a legal group key


def test_partial_then_final_equals_complete():
    rng = np.random.default_rng(9)
    n = 5000
    t = pa.table({"k": rng.integers(0, 37, n), "v": pa.array(rng.integers(-1000, 1000, n), mask=rng.random(n) < 0.1),
                  "d": rng.random(n)})
    aggs = [("sum", "v", "s"), ("avg", "d", "a"), ("count", "v", "c"), ("count_star", None, "n"), ("max", "d", "mx")]
    complete = O.hash_aggregate(t, ["k"], aggs, "complete")
    halves = [O.hash_aggregate(t.slice(0, n // 2), ["k"], aggs, "partial"), O.hash_aggregate(t.slice(n // 2), ["k"], aggs, "partial")]
    final = O.hash_aggregate(pa.concat_tables(halves), ["k"], aggs, "final")
    from util import assert_tables_equal
    assert_tables_equal(final, complete, key_cols=["k"], rtol=1e-12)


def test_sum_long_wraps_non_ansi():
    t = pa.table({"k": pa.array([0, 0], type=pa.int32()), "v": pa.array([2 ** 63 - 1, 1], type=pa.int64())})
    out = O.hash_aggregate(t, ["k"], [("sum", "v", "s")])
    assert out.column("s").to_pylist() == [-2 ** 63]      # Sum.scala non-ANSI: Add wraps


def test_q1_wscg_restatement_matches_operator_oracle():
    from spark_b200 import tpch
    n = 200_000
    c = tpch.lineitem_q1_columns(n, seed=3)
    t = tpch.lineitem_q1_table(n, seed=3)
    L = O.lib()
    k0 = np.zeros(16, np.int8); k1 = np.zeros(16, np.int8); sums = np.zeros(16 * 5); cnt = np.zeros(16, np.int64)
    ng = L.so_q1_partial_final(c["l_quantity"].ctypes.data, c["l_extendedprice"].ctypes.data, c["l_discount"].ctypes.data,
                               c["l_tax"].ctypes.data, c["l_returnflag"].ctypes.data, c["l_linestatus"].ctypes.data,
                               c["l_shipdate"].ctypes.data, n, tpch.Q1_CUTOFF, 16, k0.ctypes.data, k1.ctypes.data,
                               sums.ctypes.data, cnt.ctypes.data)
    f = O.filter_table(t, ("le", ("col", "l_shipdate"), ("lit", tpch.Q1_CUTOFF, np.int32)))
    p = O.project(f, [("l_returnflag", ("col", "l_returnflag")), ("l_linestatus", ("col", "l_linestatus")),
                      ("l_quantity", ("col", "l_quantity")), ("l_extendedprice", ("col", "l_extendedprice")),
                      ("l_discount", ("col", "l_discount")),
                      ("disc_price", ("mul", ("col", "l_extendedprice"), ("sub", ("lit", 1.0), ("col", "l_discount")))),
                      ("charge", ("mul", ("mul", ("col", "l_extendedprice"), ("sub", ("lit", 1.0), ("col", "l_discount"))),
                                  ("add", ("lit", 1.0), ("col", "l_tax"))))])
    want = O.hash_aggregate(p, ["l_returnflag", "l_linestatus"], tpch.q1_oracle_aggs())
    assert ng == want.num_rows
    got = {(int(k0[g]), int(k1[g])): (sums[g * 5:(g + 1) * 5], cnt[g]) for g in range(ng)}
    for r in zip(*[want.column(i).to_pylist() for i in range(want.num_columns)]):
        s, c_ = got[(r[0], r[1])]
        assert c_ == r[9]
        np.testing.assert_allclose([s[0], s[1], s[2], s[3]], r[2:6], rtol=1e-9)
        np.testing.assert_allclose([s[0] / c_, s[1] / c_, s[4] / c_], r[6:9], rtol=1e-9)


def test_determine_bounds_reference_vector():
    """PartitioningSuite.scala:119-125: determineBounds of the literal candidate list with 3 partitions is (0.4, 0.7); an empty
    candidate set has no bounds."""
    assert O.determine_bounds([], 10) == []
    cands = [(0.7, 2.0), (0.1, 1.0), (0.4, 1.0), (0.3, 1.0), (0.2, 1.0), (0.5, 1.0), (1.0, 3.0)]
    assert O.determine_bounds(cands, 3) == [0.4, 0.7]


def test_outer_and_existence_join_reference_answers():
    """The oracle's join (all types, residual condition) reproduces the literal answers of OuterJoinSuite.scala:191-258 and
    ExistenceJoinSuite.scala:356-461 on their literal inputs (tests/join_fixtures.py)."""
    import join_fixtures as F
    for how, want in F.OUTER_CASES.items():
        if how == "right_outer":     # the right side is preserved: hash the right side (build_outer)
            got = O.hash_join(F.OUTER_LEFT, F.OUTER_RIGHT, ["a"], ["c"], "build_outer", F.COND_B_LT_D)
        else:
            got = O.hash_join(F.OUTER_LEFT, F.OUTER_RIGHT, ["a"], ["c"], how, F.COND_B_LT_D)
        assert F.multiset(F.rows_of(got)) == F.multiset(want), how
    for how, cond, want in F.EXIST_CASES:
        got = O.hash_join(F.EXIST_LEFT, F.EXIST_RIGHT, ["a"], ["c"], how, cond)
        assert F.multiset(F.rows_of(got)) == F.multiset(want), (how, cond)
    # existence join: one boolean per streamed row = membership in the semi join's answer
    ex = O.hash_join(F.EXIST_LEFT, F.EXIST_RIGHT, ["a"], ["c"], "existence", F.COND_B_LT_D)
    assert ex.column("exists").to_pylist() == [False, False, True, True, False, False, False, False]


def _golden_oracle(case):
    import sql_goldens as G
    name, table, keys, proj, aggs, post, want, where = case
    t = table
    if proj:
        t = O.project(t, [(c, ("col", c)) for c in table.column_names] + proj)
    out = O.hash_aggregate(t, keys or [], aggs)
    if post == "drop_keys":
        out = out.select([n for _, _, n in aggs])
    elif post is not None:
        out = O.filter_table(out, post)
    return out


def test_aggregate_answers_of_the_reference_sql_goldens():
    """group-by.sql.out / having.sql.out (tests/sql_goldens.py): the oracle's aggregate reproduces the rows the reference prints."""
    import sql_goldens as G
    for case in G.CASES:
        got = _golden_oracle(case)
        rows = list(zip(*[got.column(i).to_pylist() for i in range(got.num_columns)])) if got.num_rows else []
        want = case[6]
        assert len(rows) == len(want), (case[0], rows, want)
        for g, w in zip(G.norm(rows), G.norm(want)):
            for x, y in zip(g, w):
                if isinstance(y, float):
                    assert abs(x - y) <= 1e-15 * abs(y), (case[0], case[7], x, y)       # the golden prints 17 significant digits
                else:
                    assert x == y, (case[0], case[7], g, w)


# --- sql-tests/results/order-by-nulls-ordering.sql.out (SPARK-10747): ORDER BY with NULLS FIRST / LAST in every direction, ties in
# --- arrival order, a string and a double key; and the same orders inside a window with a sliding ROWS frame -------------------------
def test_order_by_nulls_ordering_goldens():
    import sort_goldens as G
    t = G.t1()
    for orders, want in G.T1_ORDER_BY:
        got = O.sort(t, orders)
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col2", "col3")])) == want, orders
    t = G.t2()
    for orders, want in G.T2_ORDER_BY:
        got = O.sort(t, orders)
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col3", "col5")])) == want, orders


def test_order_by_nulls_ordering_window_goldens():
    import sort_goldens as G
    t = G.t1()
    for orders, want in G.T1_WINDOW:
        w = O.window(t, ["col1"], orders, [("sum", "col2", ("rows", -2, 2), 0, "sum_col2")])
        got = O.sort(w, [("sum_col2", True, True)])          # the outer ORDER BY sum_col2: ties keep the window's output order
        assert list(zip(*[got.column(c).to_pylist() for c in ("col1", "col2", "col3", "sum_col2")])) == want, orders


def test_order_by_all_goldens():
    """sql-tests/results/order-by-all.sql.out: two sort columns, every direction / NULL placement, and ORDER BY ... LIMIT 2."""
    import sort_goldens as G
    t = G.t3()
    for orders, want in G.T3_ORDER_BY:
        got = O.sort(t, orders)
        assert list(zip(got.column("g").to_pylist(), got.column("i").to_pylist())) == want, orders
    orders, k, want = G.T3_LIMIT_2
    got = O.take_ordered(t, orders, k)
    assert list(zip(got.column("g").to_pylist(), got.column("i").to_pylist())) == want


def test_not_in_null_aware_anti_join_goldens():
    """not-in-unit-tests-single-column.sql.out, uncorrelated cases 1-5: `a NOT IN (SELECT c FROM s WHERE ...)` is the null-aware anti
    join of BroadcastHashJoinExec.scala:137-162 (empty relation -> every row; a NULL in the relation -> none; NULL probe keys dropped)."""
    import join_fixtures as F
    m = pa.table({"a": pa.array([r[0] for r in F.NOT_IN_M], pa.int32()), "b": pa.array([r[1] for r in F.NOT_IN_M], pa.float64())})
    s = pa.table({"c": pa.array([r[0] for r in F.NOT_IN_S], pa.int32()), "d": pa.array([r[1] for r in F.NOT_IN_S], pa.float64())})
    for mb, (op, lit), want in F.NOT_IN_CASES:
        d = np.asarray(s.column("d"))
        sub = s.filter(pa.array(d > lit if op == ">" else d == lit))
        left = m if mb is None else m.filter(pa.array(np.asarray(m.column("b")) == mb))
        got = O.hash_join(left, sub, ["a"], ["c"], "left_anti_null_aware")
        rows = sorted(zip(got.column("a").to_pylist(), got.column("b").to_pylist()), key=lambda r: (r[0] is None, r[0] or 0))
        assert rows == sorted(want, key=lambda r: (r[0] is None, r[0] or 0)), (mb, op, lit)
