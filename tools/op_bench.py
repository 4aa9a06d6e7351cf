#!/usr/bin/env python
"""Per-operator device timings for the rows of SURVEY.md 8a (partition, aggregate, sort, join) -- achieved HBM GB/s
against the measured peak, one JSON line per case.  Shapes follow the reference's own micro-benchmarks
(sql/core/benchmarks/*-results.txt) and BASELINE.json configs.  Run on the GPU box:
    python tools/op_bench.py > gpurun_out/op_bench.jsonl
Numbers are whole-operator device time (CUDA events on the operator's stream, inputs resident in HBM, best of 5 after
2 warm-ups, inputs >> L2) plus the dominant kernel's own time from sb_profile_get."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_b200 import _capi as capi                       # noqa: E402
from spark_b200.columnar import ColumnarBatch, Stream       # noqa: E402
from spark_b200.execution import (BroadcastHashJoinExec, HashAggregateExec, HashPartitioning, LocalTableScanExec,  # noqa: E402
                                  ShuffleExchangeExec, SortExec)
from spark_b200.expressions import Sum, col                 # noqa: E402


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def timed(lib, stream, fn, kernel_names, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = None
    capi.check(lib.sb_profile_enable(1))
    for _ in range(reps):
        capi.check(lib.sb_profile_reset())
        stream.record_start()
        fn()
        stream.record_stop()
        ms = stream.elapsed_ms()
        k = {}
        for name in kernel_names:
            t, c = C.c_double(), C.c_int64()
            capi.check(lib.sb_profile_get(name.encode(), C.byref(t), C.byref(c)))
            k[name] = (t.value, c.value)
        if best is None or ms < best[0]:
            best = (ms, k)
    capi.check(lib.sb_profile_enable(0))
    return best


def emit(case, n, alg_bytes, ms, kernels, extra=None):
    pk = peak()
    line = {"case": case, "rows": n, "ms": ms, "rows_per_s": n / (ms / 1e3), "algorithmic_bytes": alg_bytes,
            "operator_gbs": alg_bytes / (ms / 1e3) / 1e9, "operator_frac_of_measured_peak": alg_bytes / (ms / 1e3) / 1e9 / pk,
            "kernels_ms": {k: v[0] for k, v in kernels.items()}, "kernel_launches": {k: v[1] for k, v in kernels.items()}}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    lib = capi.init(0)
    stream = Stream()
    rng = np.random.default_rng(0)

    if only in ("", "partition"):
        # ---- hash partition: C4-shaped fixed-width lineitem row (74 B/row), n = 2048 and 200 partitions --------------
        n = 24_000_000
        cols = {"l_orderkey": rng.integers(0, 1 << 40, n), "l_partkey": rng.integers(0, 1 << 30, n), "l_suppkey": rng.integers(0, 1 << 24, n),
                "l_quantity": rng.random(n), "l_extendedprice": rng.random(n), "l_discount": rng.random(n), "l_tax": rng.random(n),
                "l_shipdate": rng.integers(8000, 10000, n).astype(np.int32), "l_commitdate": rng.integers(8000, 10000, n).astype(np.int32),
                "l_receiptdate": rng.integers(8000, 10000, n).astype(np.int32), "l_linenumber": rng.integers(1, 8, n).astype(np.int32),
                "l_returnflag": rng.integers(65, 83, n).astype(np.int8), "l_linestatus": rng.integers(70, 80, n).astype(np.int8)}
        rowbytes = sum(a.dtype.itemsize for a in cols.values())
        batch = ColumnarBatch.from_numpy(cols, stream)
        stream.synchronize()
        for nparts in (2048, 200):
            ex = ShuffleExchangeExec(HashPartitioning(["l_orderkey"], nparts), LocalTableScanExec(batch))
            ms, k = timed(lib, stream, lambda: ex.executeColumnar(stream).close(), ["partition_rank", "partition_scatter"])
            emit("hash_partition n=%d (%d B/row)" % (nparts, rowbytes), n, 2 * rowbytes * n, ms, k)
        batch.close()

    if only in ("", "agg"):
        # ---- hash aggregate: SELECT k, SUM(v) GROUP BY k (configs[0] shape, scaled to 64M rows) ---------------------------
        n = 64_000_000
        v = rng.integers(-2 ** 31, 2 ** 31, n)
        group_list = [int(x) for x in os.environ["SB_OPBENCH_GROUPS"].split(",")] if os.environ.get("SB_OPBENCH_GROUPS") else [4, 1024, 65536, 1 << 20]
        for groups in group_list:
            b = ColumnarBatch.from_numpy({"k": rng.integers(0, groups, n), "v": v}, stream)
            stream.synchronize()
            for hint, tag in ((groups, "hinted"), (0, "no hint")):   # Spark's planner has no cardinality to pass: "no hint" is the honest case
                agg = HashAggregateExec(["k"], [(Sum(col("v")), "s")], LocalTableScanExec(b), expected_groups=hint)
                ms, k = timed(lib, stream, lambda: agg.executeColumnar(stream).close(), ["agg_update"])
                emit("hash_aggregate sum(int64) groups=%d (%s)" % (groups, tag), n, 16 * n + 16 * groups, ms, k)
            b.close()
        vd = rng.random(n) * 1e4
        for groups in (16, 1024):   # double sums accumulate through compare-and-swap in shared memory: a different regime
            b = ColumnarBatch.from_numpy({"k": rng.integers(0, groups, n), "v": vd}, stream)
            stream.synchronize()
            agg = HashAggregateExec(["k"], [(Sum(col("v")), "s")], LocalTableScanExec(b))
            ms, k = timed(lib, stream, lambda: agg.executeColumnar(stream).close(), ["agg_update"])
            emit("hash_aggregate sum(double) groups=%d (no hint)" % groups, n, 16 * n + 16 * groups, ms, k)
            b.close()

    if only in ("", "sort"):
        # ---- radix sort: 25M x 8-byte keys (SortBenchmark-results.txt:13) + 16 B records ----------------------------------
        n = 25_000_000
        b = ColumnarBatch.from_numpy({"k": rng.integers(-2 ** 63, 2 ** 63 - 1, n)}, stream)
        stream.synchronize()
        srt = SortExec([("k", True, True)], LocalTableScanExec(b))
        ms, k = timed(lib, stream, lambda: srt.executeColumnar(stream).close(), ["sort_histogram", "sort_passes", "gather"])
        emit("sort int64 keys (full range)", n, 2 * 8 * n, ms, k, {"note": "8 onesweep passes of (8 B key + 4 B row id) + key histogram + payload gather; algorithmic = one read + one write of the key column",
                                                       "pass_traffic_gbs": 8 * 24 * n / (k["sort_passes"][0] / 1e3) / 1e9 if k["sort_passes"][0] else None})
        b.close()
        b = ColumnarBatch.from_numpy({"k": rng.integers(0, 1 << 40, n)}, stream)
        stream.synchronize()
        srt = SortExec([("k", True, True)], LocalTableScanExec(b))
        ms, k = timed(lib, stream, lambda: srt.executeColumnar(stream).close(), ["sort_histogram", "sort_passes", "gather"])
        emit("sort int64 keys (40 significant bits: 3 constant bytes skipped)", n, 2 * 8 * n, ms, k)
        from spark_b200.execution import TakeOrderedAndProjectExec
        from spark_b200.expressions import SortOrder
        top = TakeOrderedAndProjectExec(10, [SortOrder("k", False)], None, LocalTableScanExec(b))
        ms, k = timed(lib, stream, lambda: top.executeColumnar(stream).close(), ["sort_passes"])
        emit("take_ordered(10) of int64 keys", n, 8 * n, ms, k, {"note": "radix select on the first sort column, then a sort of the candidates"})
        b.close()

    if only in ("", "scan"):
        # ---- scan: RLE_DICTIONARY pages (6-bit and 12-bit domains) and PLAIN doubles, decoded on the device ----------------------------
        from spark_b200 import tpch
        from spark_b200.scan import decode_chunks, encode_column
        cols = tpch.Q1_COLUMNS
        b = tpch.synth_batch("lineitem", cols, 6_000_000, 42, stream=stream)      # 24 M rows
        n = b.num_rows
        chunks = [encode_column(b, c, dictionary=c != "l_extendedprice", page_rows=1 << 19, stream=stream, pinned=True) for c in cols]
        enc = sum(ch.nbytes for ch in chunks)
        ms, k = timed(lib, stream, lambda: decode_chunks(cols, chunks, stream, b.arrow_types).close(), ["scan_decode"])
        emit("scan_decode Q1 columns (H2D of the encoded bytes included in ms)", n, enc + 38 * n, ms, k,
             {"encoded_bytes": enc, "decode_kernel_gbs": (enc + 38 * n) / (k["scan_decode"][0] / 1e3) / 1e9 if k["scan_decode"][0] else None})
        b.close()

    if only in ("", "strings"):
        # ---- string grouping keys (dictionary codes inside sb_hash_aggregate): 32 M rows, 25 and 1 M distinct strings ---------------------
        import pyarrow as pa
        n = 32_000_000
        for distinct in (25, 1_000_000):
            words = np.array(["key-%07d" % i for i in range(distinct)])
            tbl = pa.table({"s": pa.array(words[rng.integers(0, distinct, n)], type=pa.string()), "v": rng.integers(0, 1000, n)})
            b = ColumnarBatch.from_arrow(tbl, stream)
            stream.synchronize()
            agg = HashAggregateExec(["s"], [(Sum(col("v")), "t")], LocalTableScanExec(b))
            ms, k = timed(lib, stream, lambda: agg.executeColumnar(stream).close(), ["dictionary_encode", "agg_update"])
            emit("hash_aggregate sum(int64) by string key (11 B), %d distinct" % distinct, n, n * (11 + 4 + 8), ms, k)
            b.close()

    if only in ("", "window"):
        # ---- WindowExec: 16 M rows, 1 M partitions, ORDER BY one int column: row_number + running sum + lag -----------------------------
        from spark_b200.execution import WindowExec, WindowFunction
        n = 16_000_000
        b = ColumnarBatch.from_numpy({"p": rng.integers(0, 1_000_000, n), "o": rng.integers(0, 1 << 20, n).astype(np.int32), "v": rng.integers(0, 1000, n)}, stream)
        stream.synchronize()
        w = WindowExec([WindowFunction("row_number", None, None, 0, "rn"), WindowFunction("sum", "v", ("rows", None, 0), 0, "run"),
                        WindowFunction("lag", "v", None, 1, "prev")], ["p"], [("o", True, True)], LocalTableScanExec(b))
        ms, k = timed(lib, stream, lambda: w.executeColumnar(stream).close(), ["sort_passes", "gather"])
        emit("window row_number + running sum + lag, 1 M partitions", n, n * (20 + 20 + 4 + 8 + 8), ms, k)
        b.close()

    if only in ("", "decimal"):
        # ---- decimal(12, 2) SUM + AVG by 1 K groups, 32 M rows (limb sums around the aggregate) ---------------------------------------------
        import pyarrow as pa
        from spark_b200.expressions import Average
        n = 32_000_000
        raw = np.zeros((n, 2), np.int64)
        raw[:, 0] = rng.integers(0, 10 ** 11, n)
        dec = pa.Array.from_buffers(pa.decimal128(12, 2), n, [None, pa.py_buffer(raw.tobytes())])
        b = ColumnarBatch.from_arrow(pa.table({"k": rng.integers(0, 1000, n).astype(np.int32), "d": dec}), stream)
        stream.synchronize()
        agg = HashAggregateExec(["k"], [(Sum(col("d")), "s"), (Average(col("d")), "a")], LocalTableScanExec(b))
        ms, k = timed(lib, stream, lambda: agg.executeColumnar(stream).close(), ["agg_update"])
        emit("hash_aggregate sum + avg(decimal(12, 2)) groups=1000", n, n * 12, ms, k)
        b.close()

    if only in ("", "join"):
        # ---- hash join: 21M probe x 65k build (JoinBenchmark-results.txt:10) and a 16M-row build -----------------------------
        for nb, npr in ((65536, 21_000_000), (16_000_000, 64_000_000)):
            build = ColumnarBatch.from_numpy({"id": rng.permutation(nb).astype(np.int64), "payload": np.arange(nb, dtype=np.int64)}, stream)
            probe = ColumnarBatch.from_numpy({"fk": rng.integers(0, nb, npr), "x": np.arange(npr, dtype=np.int64)}, stream)
            stream.synchronize()
            j = BroadcastHashJoinExec(["fk"], ["id"], "inner", "right", LocalTableScanExec(probe), LocalTableScanExec(build))
            ms, k = timed(lib, stream, lambda: j.executeColumnar(stream).close(), ["join_build", "join_candidates", "join_probe", "join_fill", "gather"])
            emit("hash_join inner probe=%d build=%d" % (npr, nb), npr, nb * 16 + npr * 16 + npr * 32, ms, k)
            build.close(); probe.close()


if __name__ == "__main__":
    main()
