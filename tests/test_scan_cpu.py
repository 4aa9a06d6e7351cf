"""CPU tests of the scan boundary's host side: the Thrift page walker of the product (sb_parquet_chunk_pages, host only) and
the oracle's restatement of the page decoding, both against files written -- and read back -- by pyarrow (parquet-cpp)."""
import ctypes as C
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import parquet_oracle as PO


def _table(n, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal(n)
    return pa.table({
        "i32_dict": pa.array(rng.integers(0, 50, n).astype(np.int32)),
        "i64_plain": pa.array(rng.integers(-2 ** 60, 2 ** 60, n)),
        "i64_null": pa.array(rng.integers(0, 1000, n), mask=rng.random(n) < 0.2),
        "f64_dict": pa.array((rng.integers(0, 11, n) / 100.0)),
        "f64_null": pa.array(d, mask=rng.random(n) < 0.5),
        "f32": pa.array(rng.standard_normal(n).astype(np.float32)),
        "date": pa.array(rng.integers(8000, 10500, n).astype(np.int32)).cast(pa.date32()),
        "i8": pa.array(rng.integers(-3, 3, n).astype(np.int8), mask=rng.random(n) < 0.1),
        "flag": pa.array(rng.integers(0, 2, n).astype(bool)),
        "runs": pa.array(np.repeat(rng.integers(0, 4, (n + 99) // 100), 100)[:n].astype(np.int32)),   # long RLE runs
        "const": pa.array(np.full(n, 7, np.int64)),
    })


CASES = [dict(version="1.0", use_dictionary=True, n=20_000), dict(version="2.0", use_dictionary=True, n=20_000),
         dict(version="1.0", use_dictionary=False, n=5_001), dict(version="1.0", use_dictionary=True, n=1),
         dict(version="2.0", use_dictionary=["i32_dict", "f64_dict", "runs"], n=70_003)]


def write_case(tmp_path, case, seed=0):
    t = _table(case["n"], seed)
    path = os.path.join(tmp_path, "t.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=case["use_dictionary"], data_page_version=case["version"],
                   data_page_size=16 * 1024, row_group_size=case.get("row_group_size", case["n"]), write_statistics=True)
    return t, path


@pytest.mark.parametrize("case", CASES, ids=lambda c: "v%s-dict%s-n%d" % (c["version"], "sel" if isinstance(c["use_dictionary"], list) else c["use_dictionary"], c["n"]))
def test_page_walker_and_oracle_decoder_against_pyarrow(tmp_path, case):
    from spark_b200.scan import ParquetScanExec
    t, path = write_case(str(tmp_path), case)
    scan = ParquetScanExec(path)                        # footer via pyarrow, page headers via sb_parquet_chunk_pages (no device needed)
    back = pq.read_table(path)
    for name, ch in zip(scan.columns, scan.row_group_chunks(0)):
        pages = [(ch.pages[i].encoding, ch.pages[i].num_values, ch.pages[i].values_offset, ch.pages[i].values_bytes,
                  ch.pages[i].def_offset, ch.pages[i].def_bytes) for i in range(ch.npages)]
        assert sum(p[1] for p in pages) == case["n"]
        vals, valid = PO.decode_column_chunk(ch.data, pages, ch.dict_offset, ch.dict_count, ch.physical)
        col = back.column(name).combine_chunks()
        want_valid = np.asarray(col.is_valid())
        if valid is None:
            valid = np.ones(len(vals), bool)
        assert np.array_equal(valid, want_valid), name
        typ = col.type
        if pa.types.is_date32(typ):
            want = np.asarray(col.cast(pa.int32()).fill_null(0))
        elif pa.types.is_boolean(typ):
            want = np.asarray(col.fill_null(False)).astype(np.uint8)
        else:
            want = np.asarray(col.fill_null(0))
        got = vals.astype(want.dtype) if vals.dtype != want.dtype else vals
        assert np.array_equal(got[valid], want[valid]), name


def test_compressed_pages_are_rejected_not_decoded_on_the_cpu(tmp_path):
    from spark_b200 import _capi as capi
    from spark_b200.scan import ParquetScanExec
    t = _table(1000)
    path = os.path.join(str(tmp_path), "c.parquet")
    pq.write_table(t, path, compression="SNAPPY")
    with pytest.raises(capi.SparkB200Error):
        ParquetScanExec(path).row_group_chunks(0)


def test_malformed_chunk_is_an_error():
    from spark_b200 import _capi as capi
    lib = capi.load()
    junk = np.frombuffer(b"\xff" * 64, dtype=np.uint8)
    pages = (capi.sb_page * 4)()
    n, doff, dc = C.c_int32(), C.c_int64(), C.c_int32()
    assert lib.sb_parquet_chunk_pages(junk.ctypes.data, junk.nbytes, 0, pages, 4, C.byref(n), C.byref(doff), C.byref(dc)) != 0
