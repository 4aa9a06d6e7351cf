"""GPU parity of the scan boundary: Parquet files written by pyarrow (parquet-cpp), column chunks shipped to the GPU encoded,
decoded by sb_scan_decode, compared with what pyarrow's own reader returns -- every encoding / page version / NULL pattern of
tests/test_scan_cpu.py -- plus the write-side twin (sb_scan_encode) round trip and Q1 from encoded chunks."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from test_scan_cpu import CASES, write_case
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES + [dict(version="1.0", use_dictionary=True, n=300_000, row_group_size=100_000)],
                         ids=lambda c: "v%s-dict%s-n%d" % (c["version"], "sel" if isinstance(c["use_dictionary"], list) else c["use_dictionary"], c["n"]))
def test_scan_decode_equals_pyarrow_reader(gpu, stream, tmp_path, case):
    from spark_b200.scan import ParquetScanExec
    t, path = write_case(str(tmp_path), case, seed=case["n"])
    scan = ParquetScanExec(path)
    got = scan.collect(stream)
    want = pq.read_table(path)
    assert got.num_rows == want.num_rows == case["n"]
    for name in want.column_names:
        g, w = got.column(name).combine_chunks(), want.column(name).combine_chunks()
        assert g.null_count == w.null_count, name
        assert g.to_pylist() == w.cast(g.type).to_pylist() if not pa.types.is_floating(w.type) else \
            np.array_equal(np.asarray(g.fill_null(0)), np.asarray(w.fill_null(0))), name      # bit-exact, doubles included
    # row groups as an iterator of batches
    rows = 0
    for b in scan.batches(stream):
        rows += b.num_rows
        b.close()
    assert rows == case["n"]


@pytest.mark.parametrize("page_rows", [504 * 8, 1 << 16, 1000])
def test_encode_decode_round_trip(gpu, stream, page_rows):
    """sb_scan_encode (sorted dictionary + bit-packed runs of 504 values, or PLAIN) -> host -> sb_scan_decode is the identity,
    and the chunk it writes is what the oracle's restatement of the reference reader decodes too."""
    from oracle import parquet_oracle as PO
    from spark_b200 import tpch
    from spark_b200.scan import decode_chunks, encode_column
    n_orders = 50_000
    cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate", "l_orderkey", "l_linenumber"]
    b = tpch.synth_batch("lineitem", cols, n_orders, seed=4, stream=stream)
    plain = {"l_extendedprice", "l_orderkey"}
    chunks = [encode_column(b, c, dictionary=c not in plain, page_rows=page_rows, stream=stream) for c in cols]
    total = sum(ch.nbytes for ch in chunks)
    assert total < 0.45 * sum(tpch.synth_width(c) for c in cols) * b.num_rows      # the low-cardinality columns shrink to bits
    back = decode_chunks(cols, chunks, stream, b.arrow_types)
    assert_tables_equal(back.to_arrow(stream), b.to_arrow(stream), ordered=True)
    for name, ch in zip(cols, chunks):
        pages = [(ch.pages[i].encoding, ch.pages[i].num_values, ch.pages[i].values_offset, ch.pages[i].values_bytes, 0, 0) for i in range(ch.npages)]
        vals, valid = PO.decode_column_chunk(ch.data, pages, ch.dict_offset, ch.dict_count, ch.physical)
        want, _ = b.column_to_numpy(b.column_index(name), stream)
        assert valid is None and np.array_equal(vals.astype(want.dtype), want), name


def test_q1_from_encoded_chunks_streamed_into_the_aggregate(gpu, stream):
    """The e2e shape of bench.py: row groups decoded on the GPU one after the other, each folded into one aggregation state."""
    import bench
    from spark_b200 import tpch
    from spark_b200.execution import LocalTableScanExec
    from spark_b200.scan import decode_chunks, encode_column
    n_orders = 300_000
    cols = tpch.Q1_COLUMNS
    whole = tpch.synth_batch("lineitem", cols, n_orders, seed=11, stream=stream)
    n = whole.num_rows
    groups = []
    for lo in range(0, n, 400_000):
        part = whole.slice(lo, min(n, lo + 400_000), stream)
        groups.append([encode_column(part, c, dictionary=c != "l_extendedprice", page_rows=1 << 16, stream=stream) for c in cols])
        part.close()
    partial = tpch.q1_partial_plan(LocalTableScanExec(None), fused=True)
    out = partial.execute_batches((decode_chunks(cols, g, stream, whole.arrow_types) for g in groups), stream)
    got = tpch.q1_final_plan(LocalTableScanExec(out), sort=True).collect(stream)
    want = tpch.q1_final_plan(tpch.q1_partial_plan(LocalTableScanExec(whole), fused=True), sort=True).collect(stream)
    assert_tables_equal(got, want, ordered=True)
