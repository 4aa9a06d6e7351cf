"""CPU tests: the synthetic dataset (include/sb_synth.h) and the whole-stage C restatements bench.py times as the CPU
baseline (so_q1_partial_final / so_q3 / so_q5) against the operator-by-operator oracle pipelines (oracle/tpch_oracle.py,
each operator pinned by the reference's vectors in test_oracle_golden.py)."""
import ctypes as C
import datetime

import numpy as np
import pytest

from oracle import oracle as O
from oracle import tpch_oracle as TO
from spark_b200 import tpch

N_ORDERS = 30_000


def test_synth_is_a_pure_function_of_coordinates():
    whole = TO.synth_host("lineitem", tpch.SYNTH_COLUMNS["lineitem"], N_ORDERS, seed=7)
    n = tpch.synth_rows("lineitem", N_ORDERS)
    assert all(len(v) == n for v in whole.values())
    lo, cnt = 12_345, 50_001
    part = TO.synth_host("lineitem", tpch.SYNTH_COLUMNS["lineitem"], N_ORDERS, seed=7, first_row=lo, nrows=cnt)
    for c in whole:
        assert np.array_equal(whole[c][lo:lo + cnt], part[c]), c
    other = TO.synth_host("lineitem", ["l_quantity"], N_ORDERS, seed=8)
    assert not np.array_equal(other["l_quantity"], whole["l_quantity"])


def test_synth_follows_the_tpch_shape():
    li = TO.synth_host("lineitem", tpch.SYNTH_COLUMNS["lineitem"], N_ORDERS, seed=42)
    od = TO.synth_host("orders", tpch.SYNTH_COLUMNS["orders"], N_ORDERS, seed=42)
    n = len(li["l_orderkey"])
    assert n == tpch.synth_rows("lineitem", N_ORDERS) and abs(n / N_ORDERS - 4.0) < 0.01
    # every lineitem references an order; line numbers count 1..k per order; 1..7 lines per order
    keys, counts = np.unique(li["l_orderkey"], return_counts=True)
    assert np.array_equal(keys, np.sort(od["o_orderkey"])) and counts.min() == 1 and counts.max() == 7
    assert li["l_linenumber"].min() == 1 and li["l_linenumber"].max() == 7
    odate = dict(zip(od["o_orderkey"].tolist(), od["o_orderdate"].tolist()))
    d = li["l_shipdate"] - np.array([odate[k] for k in li["l_orderkey"].tolist()])
    assert d.min() >= 1 and d.max() <= 121
    assert np.all((li["l_receiptdate"] - li["l_shipdate"] >= 1) & (li["l_receiptdate"] - li["l_shipdate"] <= 30))
    assert li["l_quantity"].min() == 1 and li["l_quantity"].max() == 50
    assert set(np.unique(li["l_discount"] * 100).round().astype(int)) == set(range(11))
    assert set(np.unique(li["l_tax"] * 100).round().astype(int)) == set(range(9))
    flags = set(zip(li["l_returnflag"].tolist(), li["l_linestatus"].tolist()))
    assert flags == {(ord("A"), ord("F")), (ord("R"), ord("F")), (ord("N"), ord("F")), (ord("N"), ord("O"))}
    cust = od["o_custkey"]
    assert cust.min() >= 1 and cust.max() <= tpch.synth_rows("customer", N_ORDERS) and not np.any((cust % 3 == 0) & (cust > 1))
    # the 74 B/row fixed-width row of BASELINE.json configs[3]
    assert sum(tpch.synth_width(c) for c in tpch.CONFIG4_COLUMNS) == tpch.CONFIG4_BYTES_PER_ROW == 74


def _host(tables):
    from bench import HostData
    return HostData(N_ORDERS, 42, tables)


def test_q1_whole_stage_restatement_equals_operator_oracle():
    from bench import CPU_TABLES, cpu_q1
    _, rows = cpu_q1(_host(CPU_TABLES["q1"]))
    t = TO.synth_arrow("lineitem", tpch.Q1_COLUMNS, N_ORDERS, 42)
    want = TO.q1(t, tpch.Q1_CUTOFF)
    assert len(rows) == want.num_rows == 4
    w = {n: want.column(n).to_pylist() for n in want.column_names}
    for i, (k0, k1, sums, cnt) in enumerate(rows):       # both sorted by (flag, status)
        assert (k0, k1, cnt) == (w["l_returnflag"][i], w["l_linestatus"][i], w["count_order"][i])
        for got, name in zip(sums[:4], ["sum_qty", "sum_base_price", "sum_disc_price", "sum_charge"]):
            assert got == pytest.approx(w[name][i], rel=1e-9)
        assert sums[4] / cnt == pytest.approx(w["avg_disc"][i], rel=1e-9)


def test_q3_whole_stage_restatement_equals_operator_oracle():
    from bench import CPU_TABLES, cpu_q3
    _, rows = cpu_q3(_host(CPU_TABLES["q3"]))
    cust = TO.synth_arrow("customer", ["c_custkey", "c_mktsegment"], N_ORDERS, 42)
    orders = TO.synth_arrow("orders", tpch.SYNTH_COLUMNS["orders"], N_ORDERS, 42)
    li = TO.synth_arrow("lineitem", ["l_orderkey", "l_extendedprice", "l_discount", "l_shipdate"], N_ORDERS, 42)
    top, agg = TO.q3(cust, orders, li, tpch.Q3_SEGMENT, tpch.Q3_DATE, 10)
    assert agg.num_rows > 100 and top.num_rows == 10 == len(rows)
    epoch = datetime.date(1970, 1, 1)
    for got, i in zip(rows, range(10)):
        assert got[0] == top.column("l_orderkey")[i].as_py()
        assert got[1] == pytest.approx(top.column("revenue")[i].as_py(), rel=1e-9)
        assert got[2] == (top.column("o_orderdate")[i].as_py() - epoch).days and got[3] == top.column("o_shippriority")[i].as_py()


def test_q5_whole_stage_restatement_equals_operator_oracle():
    from bench import CPU_TABLES, cpu_q5
    _, rows = cpu_q5(_host(CPU_TABLES["q5"]))
    cust = TO.synth_arrow("customer", ["c_custkey", "c_nationkey"], N_ORDERS, 42)
    orders = TO.synth_arrow("orders", ["o_orderkey", "o_custkey", "o_orderdate"], N_ORDERS, 42)
    li = TO.synth_arrow("lineitem", ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"], N_ORDERS, 42)
    supp = TO.synth_arrow("supplier", tpch.SYNTH_COLUMNS["supplier"], N_ORDERS, 42)
    want = TO.q5(cust, orders, li, supp, tpch.nation_table(), tpch.region_table(), tpch.Q5_REGION, tpch.Q5_DATE_LO, tpch.Q5_DATE_HI)
    assert want.num_rows == len(rows) == 5
    for got, i in zip(rows, range(want.num_rows)):
        assert got[0] == want.column("n_name")[i].as_py()
        assert got[1] == pytest.approx(want.column("revenue")[i].as_py(), rel=1e-9)


def test_ranks_of_one_host_get_disjoint_cpu_blocks():
    """bench.py --gpus N: every rank binds itself to its own block of the host CPUs before the OpenMP runtime loads (with a shared
    mask OMP_PROC_BIND put all ranks' host threads on one core: 18 ms per Q1 step at N = 4 instead of 5)."""
    from bench import rank_cpu_block
    cpus = list(range(3, 131))                       # 128 CPUs, not starting at 0
    for world in (2, 4, 8):
        blocks = [rank_cpu_block(cpus, r, world) for r in range(world)]
        assert all(len(b) == 128 // world for b in blocks)
        flat = [c for b in blocks for c in b]
        assert len(set(flat)) == len(flat) and set(flat) <= set(cpus)
        assert all(b == sorted(b) and b[-1] - b[0] == len(b) - 1 for b in blocks)      # contiguous
    assert rank_cpu_block([0, 1, 2], 1, 8) == []     # fewer CPUs than ranks: leave the mask alone
    assert rank_cpu_block({5, 4, 7, 6}, 1, 2) == [6, 7]
