"""GPU parity, round 2: the synthetic dataset on the device, aggregation state across an iterator of batches, the one-rank
communicator (self exchange through the same comm.cu code the multi-GPU path runs), and the queries bench.py measures on
the dataset bench.py measures them on -- checked against the whole-stage C restatements that are bench.py's CPU baseline."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O
from oracle import tpch_oracle as TO
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("table", ["lineitem", "orders", "customer", "supplier"])
def test_device_generator_equals_host_generator(gpu, stream, table):
    from spark_b200 import tpch
    n_orders = 70_003
    cols = tpch.SYNTH_COLUMNS[table]
    rows = tpch.synth_rows(table, n_orders)
    lo = rows // 3
    b = tpch.synth_batch(table, cols, n_orders, seed=9, first_row=lo, nrows=rows - lo - 5, stream=stream)
    host = TO.synth_host(table, cols, n_orders, seed=9, first_row=lo, nrows=rows - lo - 5)
    assert b.num_rows == rows - lo - 5
    for i, c in enumerate(cols):
        vals, valid = b.column_to_numpy(i, stream)
        assert valid is None and np.array_equal(vals, host[c]), c       # bit-exact, doubles included


def _q1_state_inputs(stream, n_orders=60_000, seed=3):
    from spark_b200 import tpch
    from spark_b200.columnar import ColumnarBatch
    t = TO.synth_arrow("lineitem", tpch.Q1_COLUMNS, n_orders, seed)
    return t, ColumnarBatch.from_arrow(t, stream)


@pytest.mark.parametrize("mode", ["partial", "complete"])
@pytest.mark.parametrize("nbatches", [1, 8])
def test_aggregation_state_over_batches_equals_single_shot(gpu, stream, mode, nbatches):
    """sb_hash_agg_create / update / finish: 8 batches streamed through one state (each released right after its update)
    give what one sb_hash_aggregate over the concatenated partition gives (TungstenAggregationIterator.processInputs)."""
    from spark_b200 import tpch
    from spark_b200.execution import HashAggregateExec, LocalTableScanExec
    from spark_b200.expressions import Literal, col
    t, whole = _q1_state_inputs(stream)
    cond = col("l_shipdate") <= Literal(tpch.Q1_CUTOFF)
    agg = HashAggregateExec(tpch.Q1_KEYS, tpch.q1_aggregates(), LocalTableScanExec(whole), mode=mode, condition=cond)
    want = agg.collect(stream)
    n = whole.num_rows
    bounds = np.linspace(0, n, nbatches + 1).astype(np.int64)
    batches = (whole.slice(int(bounds[i]), int(bounds[i + 1]), stream) for i in range(nbatches))
    got = agg.execute_batches(batches, stream).to_arrow(stream)
    assert_tables_equal(got, want, key_cols=tpch.Q1_KEYS)
    if mode == "complete":
        assert_tables_equal(got, TO.q1(t, tpch.Q1_CUTOFF, sort=False), key_cols=tpch.Q1_KEYS)


def test_aggregation_state_many_groups_nulls_and_merge(gpu, stream):
    """High-cardinality keys (the parked Partial tables get compacted on the way), NULL keys / inputs, min/max/avg/count, an
    empty batch, and sb_hash_agg_merge of another state's Partial output."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashAggregateExec, LocalTableScanExec
    from spark_b200.expressions import Average, Count, Max, Min, Sum, col
    rng = np.random.default_rng(5)
    n = 900_000
    t = pa.table({"k": pa.array(rng.integers(0, 300_000, n), mask=rng.random(n) < 0.02),
                  "v": pa.array(rng.integers(-10 ** 9, 10 ** 9, n), mask=rng.random(n) < 0.1),
                  "d": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.1)})
    aggs = [(Sum(col("v")), "sv"), (Average(col("d")), "ad"), (Count(col("v")), "cv"), (Count(), "n"), (Min(col("v")), "mn"), (Max(col("d")), "mx")]
    oaggs = [("sum", "v", "sv"), ("avg", "d", "ad"), ("count", "v", "cv"), ("count_star", None, "n"), ("min", "v", "mn"), ("max", "d", "mx")]
    whole = ColumnarBatch.from_arrow(t, stream)
    agg = HashAggregateExec(["k"], aggs, LocalTableScanExec(whole), mode="complete")
    cuts = [0, 1, 250_000, 250_000, 600_000, n]                      # includes an empty batch
    st_a = agg.new_state(whole)
    for i in range(3):
        b = whole.slice(cuts[i], cuts[i + 1], stream); st_a.update(b, stream); b.close()
    part = HashAggregateExec(["k"], aggs, LocalTableScanExec(whole), mode="partial")
    st_b = part.new_state(whole)
    for i in range(3, 5):
        b = whole.slice(cuts[i], cuts[i + 1], stream); st_b.update(b, stream); b.close()
    pb = st_b.finish(stream)                                           # keys ++ buffers of the second half
    st_a.merge(pb, stream)
    pb.close(); st_b.close()
    got = st_a.finish(stream).to_arrow(stream)
    st_a.close()
    assert_tables_equal(got, O.hash_aggregate(t, ["k"], oaggs), key_cols=["k"])


@pytest.fixture()
def one_rank_comm(gpu):
    from spark_b200 import _capi as capi
    raw = C.create_string_buffer(capi.SB_UNIQUE_ID_BYTES)
    capi.check(gpu.sb_comm_get_unique_id(raw))
    capi.check(gpu.sb_comm_init(0, 1, raw.raw))
    yield gpu
    capi.check(gpu.sb_comm_destroy())


@pytest.mark.parametrize("path", ["peer-window", "nccl"])
def test_self_exchange_through_the_communicator(one_rank_comm, stream, path, sbconfig):
    """sb_hash_partition -> sb_all_to_all and sb_all_gather on a ONE-rank NCCL communicator: the counts all-gather, the receive
    layout, the window / send-recv data path, validity as bytes and back -- everything the N-rank exchange runs except the remote
    mapping.  The rank owns every partition, so the result must be the oracle's shuffle of the whole table."""
    from spark_b200 import _capi as capi
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashPartitioning, LocalTableScanExec, ShuffleExchangeExec
    lib = one_rank_comm
    sbconfig("exchange_nccl", 1 if path == "nccl" else 0)
    rng = np.random.default_rng(12)
    n, nparts = 300_007, 200
    t = pa.table({"k": pa.array(rng.integers(0, 50_000, n), mask=rng.random(n) < 0.03),
                  "d": pa.array(rng.integers(8000, 9000, n).astype(np.int32)).cast(pa.date32()),
                  "v": pa.array(rng.random(n), mask=rng.random(n) < 0.05),
                  "f": rng.integers(0, 3, n).astype(np.int8), "row": np.arange(n, dtype=np.int64),
                  # string columns ride on the fixed-width transport: dictionary codes in the all-to-all, lengths + arena in the
                  # all-gather (csrc/comm.cu)
                  "s": pa.array(np.array(["", "N", "R", "A", "BUILDING", "你好", "x" * 40])[rng.integers(0, 7, n)], type=pa.string(),
                                mask=rng.random(n) < 0.04)})
    batch = ColumnarBatch.from_arrow(t, stream)
    ex = ShuffleExchangeExec(HashPartitioning(["k", "d"], nparts), LocalTableScanExec(batch))
    out = ex.executeColumnar(stream)
    got = out.to_arrow(stream)
    offs = ex.partition_offsets
    assert offs[0] == 0 and offs[-1] == n
    pid = O.partition_ids(t, ["k", "d"], nparts)
    order = np.argsort(pid, kind="stable")                            # the shuffle writers keep arrival order inside a partition
    want = t.take(pa.array(order))
    assert got.column("row").to_pylist() == want.column("row").to_pylist()
    assert_tables_equal(got, want, ordered=True)
    assert np.array_equal(np.diff(offs), np.bincount(pid, minlength=nparts))
    # broadcast exchange of a small and of a large table
    for rows in (5, 100_000):
        small = batch.slice(0, rows, stream)
        h = C.c_void_p()
        capi.check(lib.sb_all_gather(small.handle, stream.handle, C.byref(h)))
        g = ColumnarBatch(h, small.names, small.arrow_types)
        assert_tables_equal(g.to_arrow(stream), t.slice(0, rows), ordered=True)


N_ORDERS = 400_000       # 1.6 M lineitem rows: past the run-time specialisation threshold (2^20 rows)


def _dataset(stream, tables, n_orders=N_ORDERS, seed=42):
    from spark_b200 import tpch
    from spark_b200.execution import LocalTableScanExec
    return {t: LocalTableScanExec(tpch.synth_batch(t, cols, n_orders, seed, stream=stream)) for t, cols in tables.items()}


def _host(tables, n_orders=N_ORDERS, seed=42):
    from bench import HostData
    return HostData(n_orders, seed, tables)


def test_bench_q1_equals_cpu_baseline(gpu, stream):
    import bench
    from spark_b200 import _capi as capi, tpch
    src = _dataset(stream, bench.CPU_TABLES["q1"])
    partial = tpch.q1_partial_plan(src["lineitem"], fused=True)
    got = tpch.q1_final_plan(partial, sort=True).collect(stream)
    p = partial.executeColumnar(stream); p.close()
    assert capi.load().sb_hash_aggregate_last_plan().decode().startswith("rtc:")     # the specialised kernels ran
    _, want = bench.cpu_q1(_host(bench.CPU_TABLES["q1"]))
    c = {n: got.column(n).to_pylist() for n in got.column_names}
    rows = [(c["l_returnflag"][i], c["l_linestatus"][i], [c["sum_qty"][i], c["sum_base_price"][i], c["sum_disc_price"][i], c["sum_charge"][i],
                                                           c["avg_disc"][i] * c["count_order"][i]], c["count_order"][i]) for i in range(got.num_rows)]
    assert bench.check_q1(rows, want), (rows, want)


def test_bench_q3_equals_cpu_baseline(gpu, stream):
    import bench, datetime
    from spark_b200 import tpch
    src = _dataset(stream, bench.CPU_TABLES["q3"])
    got = tpch.q3_plan(src["customer"], src["orders"], src["lineitem"]).collect(stream)
    _, want = bench.cpu_q3(_host(bench.CPU_TABLES["q3"]))
    c = {n: got.column(n).to_pylist() for n in got.column_names}
    rows = [(c["l_orderkey"][i], c["revenue"][i], (c["o_orderdate"][i] - datetime.date(1970, 1, 1)).days, c["o_shippriority"][i]) for i in range(got.num_rows)]
    assert len(rows) == 10 and bench.check_q3(rows, want), (rows, want)


def test_bench_q5_equals_cpu_baseline(gpu, stream):
    import bench
    from spark_b200 import tpch
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec
    src = _dataset(stream, bench.CPU_TABLES["q5"])
    nation = LocalTableScanExec(ColumnarBatch.from_arrow(tpch.nation_table(), stream))
    region = LocalTableScanExec(ColumnarBatch.from_arrow(tpch.region_table(), stream))
    got = tpch.q5_plan(src["customer"], src["orders"], src["lineitem"], src["supplier"], nation, region).collect(stream)
    _, want = bench.cpu_q5(_host(bench.CPU_TABLES["q5"]))
    rows = list(zip(got.column("n_name").to_pylist(), got.column("revenue").to_pylist()))
    assert len(rows) == 5 and bench.check_q5(rows, want), (rows, want)


def test_columnar_rule_collapse_keeps_the_answer(gpu, stream):
    """Agg(Filter(Project(Filter(Project)))) collapsed by B200ColumnarRule == the uncollapsed operators == the oracle."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import B200ColumnarRule, FilterExec, HashAggregateExec, LocalTableScanExec, ProjectExec
    from spark_b200.expressions import Count, Literal, Sum, col
    rng = np.random.default_rng(8)
    n = 200_000
    t = pa.table({"k": rng.integers(0, 7, n).astype(np.int32), "x": pa.array(rng.standard_normal(n) * 10, mask=rng.random(n) < 0.05)})
    batch = ColumnarBatch.from_arrow(t, stream)

    def plan():
        inner = ProjectExec([("k", col("k")), ("y", col("x") + Literal(1.0))], LocalTableScanExec(batch))
        mid = ProjectExec([("k", col("k")), ("rev", col("y") * Literal(3.0))], FilterExec(col("y") < Literal(5.0), inner))
        return HashAggregateExec(["k"], [(Sum(col("rev")), "s"), (Count(), "n")], FilterExec(col("rev") > Literal(-20.0), mid))
    plain = plan().collect(stream)
    fused_plan = B200ColumnarRule().preColumnarTransitions(plan())
    assert isinstance(fused_plan.child, LocalTableScanExec) and fused_plan.condition is not None
    fused = fused_plan.collect(stream)
    y = ("add", ("col", "x"), ("lit", 1.0))
    rev = ("mul", y, ("lit", 3.0))
    f = O.filter_table(t, ("and", ("gt", rev, ("lit", -20.0)), ("lt", y, ("lit", 5.0))))
    want = O.hash_aggregate(O.project(f, [("k", ("col", "k")), ("rev", rev)]), ["k"], [("sum", "rev", "s"), ("count_star", None, "n")])
    assert_tables_equal(plain, want, key_cols=["k"])
    assert_tables_equal(fused, want, key_cols=["k"])


@pytest.mark.parametrize("asc,nulls_first", [(True, True), (False, False)])
def test_range_partitioner_sampling_and_bounds(gpu, stream, asc, nulls_first):
    """sb_range_sample + sb_range_determine_bounds vs the oracle's restatement of RangePartitioner.determineBounds on the SAME
    sample (which rows are sampled is unpinned in the reference); then the bounds drive sb_range_partition + per-range sort =
    the global sort, and the ranges are roughly balanced (PartitioningSuite.scala:127-139: max < 3 x min)."""
    from spark_b200 import _capi as capi
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, RangePartitioning, ShuffleExchangeExec, _orders_c, range_bounds
    from spark_b200.expressions import SortOrder
    lib = gpu
    rng = np.random.default_rng(21)
    n, nparts = 400_000, 16
    vals = rng.standard_normal(n) * 100
    t = pa.table({"k": pa.array(vals, mask=rng.random(n) < 0.01), "row": np.arange(n, dtype=np.int64)})
    batch = ColumnarBatch.from_arrow(t, stream)
    order = SortOrder("k", asc, nulls_first)
    # the sample and the oracle's bounds from it
    h = C.c_void_p()
    capi.check(lib.sb_range_sample(batch.handle, _orders_c(batch, [order]), 4800, 5, stream.handle, C.byref(h)))
    sample = ColumnarBatch(h, ["k", "weight"], [pa.float64(), pa.float32()]).to_arrow(stream)
    assert sample.num_rows == 4800 and abs(sample.column("weight")[0].as_py() - n / 4800) < 1e-3
    ks, ws = sample.column("k").to_pylist(), sample.column("weight").to_pylist()
    big = float("inf")

    def sort_key(x):        # NULL rank first, then the value in the order's direction
        if x is None:
            return (0 if nulls_first else 2, 0.0)
        return (1, x if asc else -x)
    want = O.determine_bounds(list(zip(ks, ws)), nparts, key=sort_key)
    bounds = range_bounds(batch, order, nparts, stream, samplePointsPerPartitionHint=100, seed=5)
    got = bounds.to_arrow(stream).column("k").to_pylist()
    assert got == want and len(got) == nparts - 1
    # the bounds split the rows into balanced ranges and range partition + local sort = global sort
    ex = ShuffleExchangeExec(RangePartitioning(order, bounds), LocalTableScanExec(batch))
    out = ex.executeColumnar(stream)
    sizes = np.diff(ex.partition_offsets)
    assert sizes.max() < 3.0 * sizes.min()


def test_top_n_selects_instead_of_sorting(gpu, stream):
    """TakeOrderedAndProject over 3 M rows (radix select on the first order, then a sort of the few candidates) == the first k
    rows of the full stable sort, for one and two sort columns, with heavy ties on the first."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import LocalTableScanExec, SortExec, TakeOrderedAndProjectExec
    from spark_b200.expressions import SortOrder
    rng = np.random.default_rng(33)
    n = 3_000_000
    t = pa.table({"rev": np.round(rng.random(n) * 1000, 1), "d": rng.integers(8000, 9000, n).astype(np.int32), "row": np.arange(n, dtype=np.int64),
                  "c": rng.integers(0, 4, n)})
    batch = ColumnarBatch.from_arrow(t, stream)
    for orders, k in (([SortOrder("rev", False), SortOrder("d", True)], 10), ([SortOrder("c", True)], 1000), ([SortOrder("rev", True)], 50_000)):
        got = TakeOrderedAndProjectExec(k, orders, None, LocalTableScanExec(batch)).collect(stream)
        full = SortExec(orders, LocalTableScanExec(batch)).executeColumnar(stream)
        want = full.slice(0, k, stream).to_arrow(stream)
        assert got.column("row").to_pylist() == want.column("row").to_pylist()


def test_exchange_statistics_and_coalesced_read(one_rank_comm, stream):
    """MapOutputStatistics of an exchange (sb_map_output_statistics over the communicator) feed CoalesceShufflePartitions; the
    AQE read hands out one batch per CoalescedPartitionSpec and together they are the exchange's rows."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import (AQEShuffleReadExec, HashPartitioning, LocalTableScanExec, ShuffleExchangeExec, coalesce_shuffle_partitions)
    rng = np.random.default_rng(4)
    n, nparts = 200_000, 200
    t = pa.table({"k": rng.integers(0, 10_000, n), "v": rng.random(n), "row": np.arange(n, dtype=np.int64)})
    batch = ColumnarBatch.from_arrow(t, stream)
    ex = ShuffleExchangeExec(HashPartitioning(["k"], nparts), LocalTableScanExec(batch))
    stats = ex.mapOutputStatistics(batch, stream)
    pid = O.partition_ids(t, ["k"], nparts)
    assert np.array_equal(stats, np.bincount(pid, minlength=nparts) * 24)
    specs = coalesce_shuffle_partitions([stats], advisoryTargetSize=400_000, minPartitionSize=1000)
    want = O.coalesce_partitions([stats.tolist()], 400_000, 1, 1000)
    assert len(specs) == 1 and [(s.startReducerIndex, s.endReducerIndex, s.dataSize) for s in specs[0]] == want[0] and 1 < len(want[0]) < nparts
    rows = []
    for b in AQEShuffleReadExec(ex, specs[0]).batches(stream):
        rows += b.to_arrow(stream).column("row").to_pylist()
        b.close()
    assert sorted(rows) == list(range(n))


@pytest.mark.parametrize("nparts", [1, 7, 200, 256, 2048])
def test_fused_exchange_on_a_one_rank_communicator(one_rank_comm, stream, nparts):
    """sb_shuffle_exchange (the multisplit's stores go straight into the receive window, then one local split) on a one-rank NCCL
    communicator: the window, the REMOTE scatter kernel, the barrier and the receiver-side split all run; the rank owns every
    partition, so the result is the oracle's shuffle: partition-contiguous, arrival order kept inside a partition, NULLs intact."""
    from spark_b200.columnar import ColumnarBatch
    from spark_b200.execution import HashPartitioning, LocalTableScanExec, ShuffleExchangeExec
    rng = np.random.default_rng(nparts)
    n = 250_013
    t = pa.table({"k": pa.array(rng.integers(0, 60_000, n), mask=rng.random(n) < 0.03),
                  "d": pa.array(rng.integers(8000, 9000, n).astype(np.int32)).cast(pa.date32()),
                  "v": pa.array(rng.random(n), mask=rng.random(n) < 0.05),
                  "f": rng.integers(0, 3, n).astype(np.int8), "w": rng.integers(0, 1000, n).astype(np.int16), "row": np.arange(n, dtype=np.int64)})
    batch = ColumnarBatch.from_arrow(t, stream)
    ex = ShuffleExchangeExec(HashPartitioning(["k", "d"], nparts), LocalTableScanExec(batch), fused=True)
    got = ex.executeColumnar(stream).to_arrow(stream)
    offs = ex.partition_offsets
    pid = O.partition_ids(t, ["k", "d"], nparts)
    want = t.take(pa.array(np.argsort(pid, kind="stable")))
    assert offs[0] == 0 and offs[-1] == n and np.array_equal(np.diff(offs), np.bincount(pid, minlength=nparts))
    assert got.column("row").to_pylist() == want.column("row").to_pylist()
    assert_tables_equal(got, want, ordered=True)
