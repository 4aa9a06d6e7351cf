"""GPU parity: string grouping / join / sort keys (order-preserving dictionary codes, csrc/strings.cu) vs the oracle run on the
oracle's own codes, plus the reference's byte-order vectors (UTF8StringSuite.java:100-112)."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from test_strings_cpu import BINARY_COMPARE_VECTORS
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def _words(rng, n, distinct, nulls=0.0):
    alphabet = ["", "a", "ab", "abc", "abd", "b", "ba", "Z", "z", "~", "你好", "世界", "你好123", "你好122", "\x7f", "é", "abcabcabc", "abcabcabC"]
    pool = list(alphabet)
    while len(pool) < distinct:
        k = int(rng.integers(1, 24))
        pool.append("".join(chr(int(c)) for c in rng.integers(32, 0x4ff, k)))
    pool = pool[:max(distinct, 1)]
    vals = [pool[int(i)] for i in rng.integers(0, len(pool), n)]
    if nulls:
        m = rng.random(n) < nulls
        vals = [None if m[i] else v for i, v in enumerate(vals)]
    return pa.array(vals, type=pa.string())


def _encode(batch, name, stream):
    from spark_b200 import _capi as capi
    from spark_b200.columnar import ColumnarBatch
    lib = capi.load()
    codes, dic = C.c_void_p(), C.c_void_p()
    capi.check(lib.sb_dictionary_encode(batch.handle, batch.column_index(name), stream.handle if stream else None, C.byref(codes), C.byref(dic)))
    return ColumnarBatch(codes, ["code"], [pa.int32()]), ColumnarBatch(dic, ["value"], [pa.string()])


@pytest.mark.parametrize("n,distinct", [(0, 1), (1, 1), (1000, 18), (200_000, 5000), (300_000, 250_000)])
def test_dictionary_codes_preserve_equality_and_byte_order(gpu, stream, n, distinct):
    from spark_b200.columnar import ColumnarBatch
    rng = np.random.default_rng(n + distinct)
    col = _words(rng, n, distinct, nulls=0.05 if n > 1 else 0.0)
    t = pa.table({"s": col})
    b = ColumnarBatch.from_arrow(t, stream)
    codes, dic = _encode(b, "s", stream)
    got_codes, got_dict = codes.to_arrow(stream).column(0), dic.to_arrow(stream).column(0).to_pylist()
    want_codes, want_dict = O.string_codes(t.column("s"))
    assert [d.encode() for d in got_dict] == want_dict              # distinct values in UTF8String.binaryCompare order
    assert got_codes.to_pylist() == want_codes.to_pylist()         # NULL stays NULL
    codes.close(); dic.close(); b.close()


def test_reference_byte_order_vectors(gpu, stream):
    from spark_b200.columnar import ColumnarBatch
    words = sorted({w for a, b, _ in BINARY_COMPARE_VECTORS for w in (a, b)})
    b = ColumnarBatch.from_arrow(pa.table({"s": pa.array(words, type=pa.string())}), stream)
    codes, dic = _encode(b, "s", stream)
    code = dict(zip(words, codes.to_arrow(stream).column(0).to_pylist()))
    for x, y, want in BINARY_COMPARE_VECTORS:
        assert ((code[x] > code[y]) - (code[x] < code[y])) == want, (x, y)
    codes.close(); dic.close(); b.close()


@pytest.mark.parametrize("mode", ["complete", "partial_final"])
def test_group_by_string_keys(gpu, stream, mode):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashAggregateExec, LocalTableScanExec
    from spark_b200.expressions import Count, Max, Sum, col
    rng = np.random.default_rng(5)
    n = 400_000
    t = pa.table({"flag": _words(rng, n, 7, nulls=0.02), "k2": pa.array(rng.integers(0, 3, n), type=pa.int32()),
                  "name": _words(rng, n, 3000, nulls=0.01), "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64())})
    aggs = [(Sum(col("v")), "s"), (Count(col("v")), "c"), (Max(col("v")), "m")]
    keys = ["flag", "k2", "name"]
    scan = LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))
    if mode == "complete":
        got = HashAggregateExec(keys, aggs, scan).collect(stream)
    else:
        got = HashAggregateExec(keys, aggs, HashAggregateExec(keys, aggs, scan, mode="partial"), mode="final").collect(stream)
    enc, dicts = O.encode_string_columns(t, ["flag", "name"])
    want = O.decode_string_columns(O.hash_aggregate(enc, keys, [("sum", "v", "s"), ("count", "v", "c"), ("max", "v", "m")]), dicts)
    assert got.num_rows == want.num_rows
    assert_tables_equal(got, want, key_cols=keys)


@pytest.mark.parametrize("asc,nulls_first", [(True, True), (True, False), (False, False), (False, True)])
def test_sort_by_string_keys(gpu, stream, asc, nulls_first):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, SortExec
    rng = np.random.default_rng(6)
    n = 150_000
    t = pa.table({"s": _words(rng, n, 2000, nulls=0.03), "x": pa.array(rng.integers(0, 5, n), type=pa.int32()), "row": np.arange(n, dtype=np.int64)})
    orders = [("s", asc, nulls_first), ("x", True, True)]
    got = SortExec(orders, LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))).collect(stream)
    enc, dicts = O.encode_string_columns(t, ["s"])
    want = O.decode_string_columns(O.sort(enc, orders), dicts)
    assert got.column("row").to_pylist() == want.column("row").to_pylist()      # the sort is stable: one answer
    assert got.column("s").to_pylist() == want.column("s").to_pylist()


@pytest.mark.parametrize("how", ["inner", "left_outer", "left_semi", "left_anti"])
def test_join_on_string_keys(gpu, stream, how):
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec
    rng = np.random.default_rng(7)
    nb, npr = 4000, 120_000
    build = pa.table({"name": _words(rng, nb, 3000, nulls=0.02), "payload": np.arange(nb, dtype=np.int64)})
    probe = pa.table({"who": _words(rng, npr, 5000, nulls=0.02), "id": pa.array(rng.integers(0, 10, npr), type=pa.int32()),
                      "row": np.arange(npr, dtype=np.int64)})
    lb, rb = ColumnarBatch.from_arrow(probe, stream), ColumnarBatch.from_arrow(build, stream)
    got = BroadcastHashJoinExec(["who"], ["name"], how, "right", LocalTableScanExec(lb), LocalTableScanExec(rb)).collect(stream)
    benc, dicts = O.encode_string_columns(build, ["name"])
    penc, _ = O.encode_string_columns(probe, ["who"], {"who": dicts["name"]})
    want = O.hash_join(penc, benc, ["who"], ["name"], how)
    want = O.decode_string_columns(want, {"name": dicts["name"]})
    # the streamed side's strings come back as they went in (codes of values the build side never saw are -1 in the oracle's
    # encoding, so decode the oracle's `who` from the original column by row id)
    who = probe.column("who").to_pylist()
    want = want.set_column(want.column_names.index("who"), "who", pa.array([who[r] for r in want.column("row").to_pylist()], type=pa.string()))
    assert got.num_rows == want.num_rows
    assert_tables_equal(got, want, key_cols=list(want.column_names))


def test_inner_join_on_int_and_string_keys_reference_golden(gpu, stream):
    """sql-tests/results/inner-join.sql.out: SELECT tb.* FROM ta INNER JOIN tb ON ta.a = tb.a AND ta.tag = tb.tag with
    ta = {(1,'a'),(1,'b')}, tb = {(1,'a'),(1,'a'),(1,'b'),(1,'b')} -> the four rows of tb (an int key and a string key together)."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import BroadcastHashJoinExec, LocalTableScanExec
    ta = pa.table({"a": pa.array([1, 1], type=pa.int32()), "tag": pa.array(["a", "b"], type=pa.string())})
    tb = pa.table({"a2": pa.array([1, 1, 1, 1], type=pa.int32()), "tag2": pa.array(["a", "a", "b", "b"], type=pa.string())})
    for build_side, build in (("right", tb), ("left", ta)):
        got = BroadcastHashJoinExec(["a", "tag"], ["a2", "tag2"], "inner", build_side, LocalTableScanExec(ColumnarBatch.from_arrow(ta, stream)),
                                    LocalTableScanExec(ColumnarBatch.from_arrow(tb, stream))).collect(stream)
        rows = sorted(zip(got.column("a2").to_pylist(), got.column("tag2").to_pylist()))
        assert rows == [(1, "a"), (1, "a"), (1, "b"), (1, "b")]
