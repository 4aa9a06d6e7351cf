"""GPU parity: hash build/probe joins vs the oracle and the reference's join fixtures."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def _join(left, right, lk, rk, how, stream, cls="bhj"):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec, SortMergeJoinExec
    lb, rb = ColumnarBatch.from_arrow(left, stream), ColumnarBatch.from_arrow(right, stream)
    if cls == "bhj":
        plan = BroadcastHashJoinExec(lk, rk, how, "right", LocalTableScanExec(lb), LocalTableScanExec(rb))
    else:
        plan = SortMergeJoinExec(lk, rk, how, LocalTableScanExec(lb), LocalTableScanExec(rb))
    return plan.collect(stream)


def _fixtures():   # InnerJoinSuite.scala:40-64 (letters as codes), NULL keys on both sides
    upper = pa.table({"N": pa.array([1, 2, 3, 4, 5, 6, None], type=pa.int32()), "L": pa.array([0, 1, 2, 3, 4, 5, 6], type=pa.int32())})
    lower = pa.table({"n": pa.array([1, 2, 3, 4, None], type=pa.int32()), "l": pa.array([10, 11, 12, 13, 14], type=pa.int32())})
    return upper, lower


@pytest.mark.parametrize("cls", ["bhj", "smj"])
@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti"])
def test_reference_join_fixture(gpu, stream, how, cls):
    upper, lower = _fixtures()
    got = _join(upper, lower, ["N"], ["n"], how, stream, cls)
    want = O.hash_join(upper, lower, ["N"], ["n"], how)
    assert_tables_equal(got, want, key_cols=list(want.column_names))
    if how == "inner":
        rows = sorted(zip(*[got.column(i).to_pylist() for i in range(4)]))
        assert rows == [(1, 0, 1, 10), (2, 1, 2, 11), (3, 2, 3, 12), (4, 3, 4, 13)]     # InnerJoinSuite.scala:164-171


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti"])
def test_random_join_with_duplicates_and_nulls(gpu, stream, how):
    rng = np.random.default_rng(5)
    nl, nr = 30000, 5000
    left = pa.table({"k": pa.array(rng.integers(0, 4000, nl), mask=rng.random(nl) < 0.05), "lv": rng.random(nl),
                     "ls": pa.array(["L%d" % (i % 13) for i in range(nl)])})
    right = pa.table({"k2": pa.array(rng.integers(0, 4000, nr), mask=rng.random(nr) < 0.05),
                      "rv": pa.array(rng.integers(0, 100, nr).astype(np.int32), mask=rng.random(nr) < 0.1)})
    got = _join(left, right, ["k"], ["k2"], how, stream)
    want = O.hash_join(left, right, ["k"], ["k2"], how)
    assert_tables_equal(got, want, key_cols=[c for c in want.column_names if c != "lv"])


def test_multi_column_packed_keys(gpu, stream):
    # HashJoin.rewriteKeyExpr: integral keys totalling <= 8 bytes are packed into one long
    rng = np.random.default_rng(6)
    nl, nr = 20000, 3000
    left = pa.table({"a": rng.integers(-50, 50, nl).astype(np.int32), "b": rng.integers(-3, 3, nl).astype(np.int16),
                     "c": rng.integers(0, 2, nl).astype(np.int8), "x": np.arange(nl)})
    right = pa.table({"a2": rng.integers(-50, 50, nr).astype(np.int32), "b2": rng.integers(-3, 3, nr).astype(np.int16),
                      "c2": rng.integers(0, 2, nr).astype(np.int8), "y": np.arange(nr)})
    got = _join(left, right, ["a", "b", "c"], ["a2", "b2", "c2"], "inner", stream)
    want = O.hash_join(left, right, ["a", "b", "c"], ["a2", "b2", "c2"], "inner")
    assert_tables_equal(got, want, key_cols=["x", "y"])


def test_empty_sides(gpu, stream):
    left = pa.table({"k": pa.array([1, 2, 3], type=pa.int64()), "v": pa.array([1.0, 2.0, 3.0])})
    right = pa.table({"k2": pa.array([], type=pa.int64()), "w": pa.array([], type=pa.int32())})
    assert _join(left, right, ["k"], ["k2"], "inner", stream).num_rows == 0
    assert _join(left, right, ["k"], ["k2"], "left_anti", stream).num_rows == 3
    lo = _join(left, right, ["k"], ["k2"], "left_outer", stream)
    assert lo.num_rows == 3 and lo.column("w").null_count == 3
    assert _join(left.slice(0, 0), left.rename_columns(["k2", "w"]), ["k"], ["k2"], "inner", stream).num_rows == 0


def test_pk_fk_join_full_size_properties(gpu, stream):
    """JoinBenchmark shape (21M probe x 65k build, long key): every probe row finds exactly its key."""
    rng = np.random.default_rng(8)
    nb, npr = 65536, 21_000_000
    build = pa.table({"id": rng.permutation(nb).astype(np.int64), "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"fk": rng.integers(0, 2 * nb, npr), "row": np.arange(npr, dtype=np.int64)})
    got = _join(probe, build, ["fk"], ["id"], "inner", stream)
    fk = np.asarray(got.column("fk")); idc = np.asarray(got.column("id")); row = np.asarray(got.column("row"))
    want_rows = np.nonzero(np.asarray(probe.column("fk")) < nb)[0]
    assert np.array_equal(fk, idc)
    assert np.array_equal(row, want_rows)                                   # streamed order is preserved
    inv = np.empty(nb, np.int64); inv[np.asarray(build.column("id"))] = np.arange(nb)
    assert np.array_equal(np.asarray(got.column("payload")), inv[fk])


def _plan_join(left, right, lk, rk, how, stream, build_side="right", condition=None, cls=None):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, ShuffledHashJoinExec
    lb, rb = ColumnarBatch.from_arrow(left, stream), ColumnarBatch.from_arrow(right, stream)
    return (cls or ShuffledHashJoinExec)(lk, rk, how, build_side, LocalTableScanExec(lb), LocalTableScanExec(rb), condition).collect(stream)


def test_outer_join_suite_fixtures(gpu, stream):
    """OuterJoinSuite.scala:191-258: left / right / full outer with the residual condition b < d, both build sides."""
    import join_fixtures as F
    from spark_b200.expressions import col
    cond = col("b") < col("d")
    for how, want in F.OUTER_CASES.items():
        sides = ["right", "left"] if how != "full_outer" else ["right", "left"]
        for side in sides:
            got = _plan_join(F.OUTER_LEFT, F.OUTER_RIGHT, ["a"], ["c"], how, stream, side, cond)
            assert got.column_names == ["a", "b", "c", "d"]
            assert F.multiset(F.rows_of(got)) == F.multiset(want), (how, side)


def test_existence_join_suite_fixtures(gpu, stream):
    """ExistenceJoinSuite.scala:356-461: left semi / anti with and without the residual condition, the existence join."""
    import join_fixtures as F
    from spark_b200.expressions import col
    for how, cond, want in F.EXIST_CASES:
        c = None if cond is None else (col("b") < col("d"))
        got = _plan_join(F.EXIST_LEFT, F.EXIST_RIGHT, ["a"], ["c"], how, stream, "right", c)
        assert F.multiset(F.rows_of(got)) == F.multiset(want), (how, cond)
    ex = _plan_join(F.EXIST_LEFT, F.EXIST_RIGHT, ["a"], ["c"], "existence", stream, "right", col("b") < col("d"))
    assert ex.column_names == ["a", "b", "exists"] and ex.column("exists").to_pylist() == [False, False, True, True, False, False, False, False]
    ex2 = _plan_join(F.EXIST_LEFT, F.EXIST_RIGHT, ["a"], ["c"], "existence", stream, "right", None)
    assert ex2.column("exists").to_pylist() == [False, False, True, True, True, False, False, True]


@pytest.mark.parametrize("how", ["full_outer", "right_outer", "existence", "left_anti_null_aware"])
@pytest.mark.parametrize("with_condition", [False, True])
def test_random_build_preserving_joins(gpu, stream, how, with_condition):
    from spark_b200.expressions import col
    if how == "left_anti_null_aware" and with_condition:
        pytest.skip("NAAJ takes no residual condition")
    rng = np.random.default_rng(15)
    nl, nr = 40_000, 9_000
    left = pa.table({"k": pa.array(rng.integers(0, 6000, nl), mask=rng.random(nl) < 0.03), "lv": rng.integers(0, 100, nl)})
    right = pa.table({"k2": pa.array(rng.integers(0, 6000, nr), mask=(rng.random(nr) < 0.03) if how != "left_anti_null_aware" else None),
                      "rv": rng.integers(0, 100, nr)})
    cond = (col("lv") < col("rv")) if with_condition else None
    ocond = ("lt", ("col", "lv"), ("col", "rv")) if with_condition else None
    got = _plan_join(left, right, ["k"], ["k2"], how, stream, "right", cond)
    want = O.hash_join(left, right, ["k"], ["k2"], "build_outer" if how == "right_outer" else how, ocond)
    assert_tables_equal(got, want, key_cols=list(want.column_names))
    if how == "left_anti_null_aware":      # a NULL key in the relation empties the answer; an empty relation keeps every row
        right_null = pa.table({"k2": pa.array([1, None], type=pa.int64()), "rv": pa.array([1, 2], type=pa.int64())})
        assert _plan_join(left, right_null, ["k"], ["k2"], how, stream).num_rows == 0
        assert _plan_join(left, right.slice(0, 0), ["k"], ["k2"], how, stream).num_rows == nl


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti", "left_anti_null_aware"])
@pytest.mark.parametrize("keys", ["dense", "mid", "sorted", "sparse", "negative", "nullable"])
@pytest.mark.parametrize("unique", [False, True])
@pytest.mark.parametrize("variant", [0, 1])
def test_long_streamed_side_goes_through_the_candidate_pass(gpu, stream, how, keys, unique, variant, sbconfig):
    """>= 2^20 streamed rows: one fused pass (pushed-down filter + key + prefilter) marks the candidates.  `dense` keys make the
    prefilter an exact key-range bitmap, `sparse` (60-bit) keys a Bloom filter; `negative` crosses zero; `nullable` takes the
    general key path; `mid` is a key range too sparse for the direct-address table but dense enough for the bitmap; `sorted` are
    such keys arriving in ascending order (no slot table: the row of a key is its rank among the bitmap's set bits).  With `unique`
    build keys an exact prefilter settles semi / anti joins in the pass and inner / outer joins take one lookup per candidate (a
    dense range through row_of[key - min]); duplicates go through count / scan / fill.  The streamed row count is not a multiple of 16 (ragged tail), the filter is either a conjunction of
    comparisons (evaluated inside the pass) or an OR (evaluated to a mask first)."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, FilterExec, LocalTableScanExec
    from spark_b200.expressions import col, lit
    if variant == 1 and (how in ("left_semi", "left_anti") or not unique):
        pytest.skip("the lane-strided pass (join_cand = 1) is covered by the other join types")
    sbconfig("join_cand", variant)
    rng = np.random.default_rng(11)
    nb, npr = 50_000, (1 << 20) + 12345
    if keys == "sparse":
        universe = rng.integers(0, 1 << 60, 4 * nb)
    elif keys == "negative":
        universe = np.arange(-2 * nb, 2 * nb, dtype=np.int64)
    elif keys in ("mid", "sorted"):
        universe = np.arange(0, 20 * nb, dtype=np.int64) * 3 + 17
    else:
        universe = np.arange(1000, 1000 + 4 * nb, dtype=np.int64)
    bk = rng.choice(universe, nb, replace=False)
    if keys == "sorted":                                       # strictly ascending build keys: the row of a key is its rank in the bitmap
        if not unique:
            pytest.skip("sorted mode needs unique keys")
        bk = np.sort(bk)
    if not unique:
        bk[:100] = bk[100:200]                                 # duplicate build keys: the count / fill passes still see them
    pk = rng.choice(universe, npr)
    pmask = rng.random(npr) < 0.03 if keys == "nullable" else None
    build = pa.table({"id": pa.array(bk, type=pa.int64()), "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"fk": pa.array(pk, type=pa.int64(), mask=pmask), "v": rng.integers(0, 100, npr).astype(np.int32),
                      "row": np.arange(npr, dtype=np.int64)})
    conds = [(col("v") < lit(40)) & (col("row") >= lit(7))]
    if how in ("inner", "left_anti"):
        conds.append((col("v") < lit(10)) | (col("v") > lit(80)))
    for cond in conds:
        lb, rb = ColumnarBatch.from_arrow(probe, stream), ColumnarBatch.from_arrow(build, stream)
        plan = BroadcastHashJoinExec(["fk"], ["id"], how, "right", FilterExec(cond, LocalTableScanExec(lb)), LocalTableScanExec(rb))
        got = plan.collect(stream)
        want = O.hash_join(O.filter_table(probe, cond.sexpr()), build, ["fk"], ["id"], how)
        assert got.num_rows == want.num_rows
        assert_tables_equal(got, want, key_cols=list(want.column_names))
        if how in ("inner", "left_semi", "left_anti", "left_anti_null_aware"):   # streamed order is preserved
            r = np.asarray(got.column("row"))
            assert np.all(r[1:] >= r[:-1])


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti"])
@pytest.mark.parametrize("npr", [40_000, (1 << 20) + 555])
def test_sparse_sorted_relation_keeps_the_exact_bitmap(gpu, stream, how, npr):
    """Ascending build keys spread thinly over a wide range (more than 64 bits of range per key, more than 2^23 in all): still an exact
    bitmap + rank prefix (csrc/join.cu: sorted sparse keys up to a 128 MB bitmap), probed by a short side and through the candidate pass."""
    rng = np.random.default_rng(npr)
    nb = 20_000
    bk = np.sort(rng.choice(np.arange(7, 1000 * nb, dtype=np.int64), nb, replace=False))
    build = pa.table({"id": pa.array(bk, type=pa.int64()), "payload": np.arange(nb, dtype=np.int64)})
    fk = np.where(rng.random(npr) < 0.3, bk[rng.integers(0, nb, npr)], rng.integers(-5, 1000 * nb + 5, npr))
    probe = pa.table({"fk": pa.array(fk, type=pa.int64(), mask=rng.random(npr) < 0.02), "row": np.arange(npr, dtype=np.int64)})
    got = _plan_join(probe, build, ["fk"], ["id"], how, stream)
    want = O.hash_join(probe, build, ["fk"], ["id"], how)
    assert got.num_rows == want.num_rows
    assert_tables_equal(got, want, key_cols=list(want.column_names))


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti", "full_outer", "right_outer", "existence"])
def test_short_streamed_side_over_a_sorted_relation(gpu, stream, how):
    """Build keys in strictly ascending order (a key-ordered scan, the output of a join over one): the relation is a bitmap + rank
    prefix without a slot table; a short streamed side goes through the general count / fill passes, which must find rows by rank."""
    rng = np.random.default_rng(21)
    nb, npr = 30_000, 50_000
    bk = np.sort(rng.choice(np.arange(0, 40 * nb, dtype=np.int64), nb, replace=False))
    build = pa.table({"id": pa.array(bk, type=pa.int64()), "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"fk": pa.array(rng.integers(-5, 40 * nb + 5, npr), type=pa.int64(), mask=rng.random(npr) < 0.02), "row": np.arange(npr, dtype=np.int64)})
    got = _plan_join(probe, build, ["fk"], ["id"], how, stream)
    want = O.hash_join(probe, build, ["fk"], ["id"], "build_outer" if how == "right_outer" else how)   # the build side is the right side
    assert got.num_rows == want.num_rows
    assert_tables_equal(got, want, key_cols=list(want.column_names))


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti", "full_outer", "right_outer", "existence"])
@pytest.mark.parametrize("npr", [60_000, (1 << 20) + 777])
def test_join_keys_wider_than_64_bits(gpu, stream, how, npr):
    """Three key columns of 64 + 32 + 64 bits: they cannot be packed into one word (HashJoin.rewriteKeyExpr gives up too), so the
    relation keeps a hash per slot and every hash match is verified against the build row's key columns.  Keys agree on two columns
    and differ on the third, duplicates and NULLs on both sides; the long streamed side goes through the candidate pass."""
    rng = np.random.default_rng(31)
    nb = 20_000
    def side(n, nulls):
        a = rng.integers(0, 300, n)
        b = rng.integers(0, 40, n).astype(np.int32)
        c = rng.integers(-5, 5, n) * (1 << 40)
        return pa.array(a, type=pa.int64(), mask=rng.random(n) < nulls), pa.array(b, type=pa.int32()), pa.array(c, type=pa.int64(), mask=rng.random(n) < nulls)
    ba, bb, bc = side(nb, 0.01)
    pa_, pb, pc = side(npr, 0.01)
    build = pa.table({"a2": ba, "b2": bb, "c2": bc, "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"a": pa_, "b": pb, "c": pc, "row": np.arange(npr, dtype=np.int64)})
    got = _plan_join(probe, build, ["a", "b", "c"], ["a2", "b2", "c2"], how, stream)
    want = O.hash_join(probe, build, ["a", "b", "c"], ["a2", "b2", "c2"], "build_outer" if how == "right_outer" else how)
    assert got.num_rows == want.num_rows
    assert_tables_equal(got, want, key_cols=list(want.column_names))


# Runtime filters (InjectRuntimeFilter.scala:47-100: `applicationKey IN (creation side keys)` as a might-contain test below the join).
# Exact creation sides (dense integer keys -> bitmap) make the filter an exact semi join: the result must equal the oracle's join of
# the semi-joined streamed side.  Sparse creation keys take the Bloom filter, which may let non-members through: the result is
# sandwiched between the filtered and the unfiltered join.
@pytest.mark.parametrize("how", ["inner", "left_semi"])
@pytest.mark.parametrize("n", [5000, (1 << 20) + 4097])        # byte-mask fallback / inside the candidate pass
@pytest.mark.parametrize("rf_width", [8, 4])
def test_runtime_filter_exact(gpu, stream, how, n, rf_width):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec, RuntimeFilter
    rng = np.random.default_rng(n + rf_width)
    s = rng.integers(0, 1000, n)
    s = s.astype(np.int64) if rf_width == 8 else s.astype(np.int32)
    left = pa.table({"k": rng.integers(0, 60000, n), "s": s, "x": np.arange(n, dtype=np.int64)})
    right = pa.table({"k2": rng.permutation(60000)[:20000].astype(np.int64), "rv": np.arange(20000, dtype=np.int32)})
    ck = rng.permutation(1000)[:300]
    creation = pa.table({"c": ck.astype(np.int64) if rf_width == 8 else ck.astype(np.int32)})
    lb, rb, cb = (ColumnarBatch.from_arrow(t, stream) for t in (left, right, creation))
    plan = BroadcastHashJoinExec(["k"], ["k2"], how, "right", LocalTableScanExec(lb), LocalTableScanExec(rb),
                                 runtimeFilters=[RuntimeFilter("s", "c", LocalTableScanExec(cb))])
    got = plan.collect(stream)
    keep = np.isin(np.asarray(left.column("s")), ck)
    want = O.hash_join(left.filter(pa.array(keep)), right, ["k"], ["k2"], how)
    assert_tables_equal(got, want, key_cols=list(want.column_names))


def test_runtime_filter_bloom_is_a_superset_filter(gpu, stream):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec, RuntimeFilter
    rng = np.random.default_rng(99)
    n = (1 << 20) + 77
    left = pa.table({"k": rng.integers(0, 50000, n), "s": rng.integers(0, 2 ** 62, n), "x": np.arange(n, dtype=np.int64)})
    sv = np.asarray(left.column("s"))
    right = pa.table({"k2": np.arange(50000, dtype=np.int64)})
    ck = np.concatenate([sv[rng.permutation(n)[:100000]], rng.integers(0, 2 ** 62, 50000)])     # sparse 62-bit keys: Bloom prefilter
    lb, rb, cb = (ColumnarBatch.from_arrow(t, stream) for t in (left, right, pa.table({"c": ck})))
    got = BroadcastHashJoinExec(["k"], ["k2"], "left_semi", "right", LocalTableScanExec(lb), LocalTableScanExec(rb),
                                runtimeFilters=[RuntimeFilter("s", "c", LocalTableScanExec(cb))]).collect(stream)
    gx = np.sort(np.asarray(got.column("x")))
    members = np.flatnonzero(np.isin(sv, ck))
    assert len(np.unique(gx)) == len(gx)
    assert np.all(np.isin(members, gx)), "a runtime filter must never drop a member"
    assert len(gx) < 0.3 * n, "and it should drop most non-members (%d of %d rows kept, %d members)" % (len(gx), n, len(members))


def test_runtime_filter_rejected_on_outer_joins(gpu, stream):
    from spark_b200 import _capi as capi
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec, RuntimeFilter
    t = pa.table({"k": np.arange(10, dtype=np.int64)})
    b = [ColumnarBatch.from_arrow(t, stream) for _ in range(3)]
    plan = BroadcastHashJoinExec(["k"], ["k"], "left_outer", "right", LocalTableScanExec(b[0]), LocalTableScanExec(b[1].rename(["k"])),
                                 runtimeFilters=[RuntimeFilter("k", "k", LocalTableScanExec(b[2]))])
    with pytest.raises(capi.SparkB200Error):
        plan.collect(stream)


def test_relation_built_on_one_stream_probed_on_another(gpu):
    """sb_join_build* leaves an event behind instead of draining its stream; a probe on ANOTHER stream must wait for it on the device
    (include/spark_b200.h, "Streams and ordering").  The build is long enough to still be running when the probe is enqueued."""
    from spark_b200.columnar import ColumnarBatch, Stream
    from spark_b200.execution import HashedRelation, probe_join
    rng = np.random.default_rng(1234)
    nb, npr = 4_000_000, 300_000
    bk = rng.permutation(8 * nb)[:nb].astype(np.int64)
    build = pa.table({"id": bk, "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"fk": rng.integers(0, 8 * nb, npr), "row": np.arange(npr, dtype=np.int64)})
    want = O.hash_join(probe, build, ["fk"], ["id"], "inner")
    for _ in range(3):
        s1, s2 = Stream(), Stream()
        bb = ColumnarBatch.from_arrow(build, s1)
        pb = ColumnarBatch.from_arrow(probe, s2)
        rel = HashedRelation(bb, ["id"], s1)
        out = probe_join(rel, pb, ["fk"], "inner", s2)
        got = out.to_arrow(s2)
        s1.synchronize()
        for x in (out, rel, pb, bb):
            x.close()
        assert got.num_rows == want.num_rows
        assert_tables_equal(got, want, key_cols=list(want.column_names))
