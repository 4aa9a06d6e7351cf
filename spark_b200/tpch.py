"""Deterministic TPC-H-shaped synthetic columns and the Q1/Q3/Q5 physical plans.

The reference ships the TPC-H queries (sql/core/src/test/resources/tpch/q{1,3,5}.sql), a schema
(sql/core/src/test/scala/org/apache/spark/sql/TPCHBase.scala:36-92) and plan-shape goldens
(sql/core/src/test/resources/tpch-plan-stability/) but no data generator, so value distributions follow
the TPC-H specification.  Layout decisions, stated once: money/quantity columns are DOUBLE (north_star's
1e-6 relative tolerance for floating SUM/AVG applies), keys are BIGINT, dates are DATE (date32), the
single-character flags and the market segment / nation names are dictionary codes (int8 / int32).
Only the columns a query references are generated.
"""
from __future__ import annotations

import datetime

import numpy as np
import pyarrow as pa

from .expressions import Average, Count, Literal, Sum, col
from .execution import (FilterExec, HashAggregateExec, LocalTableScanExec, ProjectExec, SortExec, SparkPlan)

EPOCH = datetime.date(1970, 1, 1)


def days(y, m, d):
    return (datetime.date(y, m, d) - EPOCH).days


LINEITEM_PER_SF = 6_001_215          # TPC-H spec row counts at SF1
ORDERS_PER_SF = 1_500_000
CUSTOMER_PER_SF = 150_000
SUPPLIER_PER_SF = 10_000

Q1_CUTOFF = days(1998, 12, 1) - 90   # date '1998-12-01' - interval '90' day  (q1.sql)
Q3_DATE = days(1995, 3, 15)
Q3_SEGMENT = 1                        # dictionary code of 'BUILDING' in SEGMENTS
SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"]
Q5_DATE_LO = days(1994, 1, 1)
Q5_DATE_HI = days(1995, 1, 1)
Q5_REGION = 2                         # 'ASIA'
CURRENT_DATE = days(1995, 6, 17)
ORDERDATE_MIN, ORDERDATE_MAX = days(1992, 1, 1), days(1998, 8, 2)

FLAG_A, FLAG_N, FLAG_R = ord("A"), ord("N"), ord("R")
STATUS_F, STATUS_O = ord("F"), ord("O")


def lineitem_q1_columns(n_rows: int, seed: int = 42, out: dict = None) -> dict:
    """The seven lineitem columns Q1 scans (38 B/row): numpy arrays, optionally written into `out` buffers."""
    rng = np.random.default_rng(seed)

    def buf(name, dtype):
        if out is not None:
            return out[name]
        return np.empty(n_rows, dtype)
    qty = buf("l_quantity", np.float64)
    price = buf("l_extendedprice", np.float64)
    disc = buf("l_discount", np.float64)
    tax = buf("l_tax", np.float64)
    rflag = buf("l_returnflag", np.int8)
    lstat = buf("l_linestatus", np.int8)
    ship = buf("l_shipdate", np.int32)
    chunk = 1 << 22
    for lo in range(0, n_rows, chunk):
        hi = min(n_rows, lo + chunk)
        m = hi - lo
        q = rng.integers(1, 51, m)
        qty[lo:hi] = q
        retail = rng.integers(90000, 200001, m)                      # part retail price in cents
        price[lo:hi] = (q * retail) / 100.0
        disc[lo:hi] = rng.integers(0, 11, m) / 100.0
        tax[lo:hi] = rng.integers(0, 9, m) / 100.0
        odate = rng.integers(ORDERDATE_MIN, ORDERDATE_MAX + 1, m)
        sdate = odate + rng.integers(1, 122, m)
        rdate = sdate + rng.integers(1, 31, m)
        ship[lo:hi] = sdate
        ra = np.where(rng.integers(0, 2, m) == 0, FLAG_R, FLAG_A)
        rflag[lo:hi] = np.where(rdate <= CURRENT_DATE, ra, FLAG_N)
        lstat[lo:hi] = np.where(sdate > CURRENT_DATE, STATUS_O, STATUS_F)
    return {"l_quantity": qty, "l_extendedprice": price, "l_discount": disc, "l_tax": tax,
            "l_returnflag": rflag, "l_linestatus": lstat, "l_shipdate": ship}


Q1_DTYPES = {"l_quantity": np.float64, "l_extendedprice": np.float64, "l_discount": np.float64, "l_tax": np.float64,
             "l_returnflag": np.int8, "l_linestatus": np.int8, "l_shipdate": np.int32}
Q1_BYTES_PER_ROW = 4 * 8 + 4 + 2     # 38 B/row


def lineitem_q1_table(n_rows: int, seed: int = 42) -> pa.Table:
    c = lineitem_q1_columns(n_rows, seed)
    return pa.table({"l_quantity": c["l_quantity"], "l_extendedprice": c["l_extendedprice"], "l_discount": c["l_discount"],
                     "l_tax": c["l_tax"], "l_returnflag": c["l_returnflag"], "l_linestatus": c["l_linestatus"],
                     "l_shipdate": pa.array(c["l_shipdate"]).cast(pa.date32())})


Q1_KEYS = ["l_returnflag", "l_linestatus"]


def q1_aggregates():
    """The eight aggregate expressions of q1.sql, in order."""
    disc_price = col("l_extendedprice") * (Literal(1) - col("l_discount"))
    charge = disc_price * (Literal(1) + col("l_tax"))
    return [(Sum(col("l_quantity")), "sum_qty"), (Sum(col("l_extendedprice")), "sum_base_price"),
            (Sum(disc_price), "sum_disc_price"), (Sum(charge), "sum_charge"),
            (Average(col("l_quantity")), "avg_qty"), (Average(col("l_extendedprice")), "avg_price"),
            (Average(col("l_discount")), "avg_disc"), (Count(), "count_order")]


def q1_partial_plan(scan: SparkPlan, fused: bool = True) -> SparkPlan:
    """Stage 1 of the golden plan (tpch-plan-stability/q1/simplified.txt): Scan -> Filter -> Project ->
    HashAggregate(partial).  fused=True hands the aggregate its child predicate and expressions directly
    (what B200ColumnarRule produces); fused=False keeps Filter and Project as separate operators."""
    cond = col("l_shipdate") <= Literal(Q1_CUTOFF)
    if fused:
        return HashAggregateExec(Q1_KEYS, q1_aggregates(), scan, mode="partial", condition=cond)
    filt = FilterExec(cond, scan)
    proj = ProjectExec(["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus"], filt)
    return HashAggregateExec(Q1_KEYS, q1_aggregates(), proj, mode="partial")


def q1_final_plan(partial: SparkPlan, sort: bool = True) -> SparkPlan:
    final = HashAggregateExec(Q1_KEYS, q1_aggregates(), partial, mode="final")
    if not sort:
        return final
    return SortExec([("l_returnflag", True, True), ("l_linestatus", True, True)], final)


def q1_oracle_aggs():
    """(func, input-expr-or-column, name) for the oracle: expressions are projected first."""
    return [("sum", "l_quantity", "sum_qty"), ("sum", "l_extendedprice", "sum_base_price"),
            ("sum", "disc_price", "sum_disc_price"), ("sum", "charge", "sum_charge"),
            ("avg", "l_quantity", "avg_qty"), ("avg", "l_extendedprice", "avg_price"),
            ("avg", "l_discount", "avg_disc"), ("count_star", None, "count_order")]


# ------------------------------------------------------------------------------------------ Q3 / Q5 tables
def customer_table(sf: float, seed: int = 7) -> pa.Table:
    n = max(1, int(CUSTOMER_PER_SF * sf))
    rng = np.random.default_rng(seed)
    return pa.table({"c_custkey": np.arange(1, n + 1, dtype=np.int64),
                     "c_mktsegment": rng.integers(0, 5, n).astype(np.int8),
                     "c_nationkey": rng.integers(0, 25, n).astype(np.int64)})


def orders_table(sf: float, seed: int = 8) -> pa.Table:
    n = max(1, int(ORDERS_PER_SF * sf))
    ncust = max(1, int(CUSTOMER_PER_SF * sf))
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.int64)
    okey = (i // 8) * 32 + (i % 8) + 1                               # sparse order keys (8 of every 32)
    cust = rng.integers(1, ncust + 1, n).astype(np.int64)
    cust = np.where(cust % 3 == 0, np.maximum(cust - 1, 1), cust)    # a third of the customers have no orders
    odate = rng.integers(ORDERDATE_MIN, ORDERDATE_MAX + 1, n).astype(np.int32)
    return pa.table({"o_orderkey": okey, "o_custkey": cust, "o_orderdate": pa.array(odate).cast(pa.date32()),
                     "o_shippriority": np.zeros(n, np.int32)})


def lineitem_join_table(orders: pa.Table, sf: float, seed: int = 9) -> pa.Table:
    """lineitem columns Q3/Q5 reference: 1..7 lines per order, ship date 1..121 days after the order date."""
    rng = np.random.default_rng(seed)
    okey = np.asarray(orders.column("o_orderkey"))
    odate = np.asarray(orders.column("o_orderdate").cast(pa.int32()))
    lines = rng.integers(1, 8, len(okey))
    l_okey = np.repeat(okey, lines)
    l_odate = np.repeat(odate, lines)
    n = len(l_okey)
    q = rng.integers(1, 51, n)
    retail = rng.integers(90000, 200001, n)
    nsupp = max(1, int(SUPPLIER_PER_SF * sf))
    return pa.table({"l_orderkey": l_okey,
                     "l_suppkey": rng.integers(1, nsupp + 1, n).astype(np.int64),
                     "l_extendedprice": (q * retail) / 100.0,
                     "l_discount": rng.integers(0, 11, n) / 100.0,
                     "l_shipdate": pa.array((l_odate + rng.integers(1, 122, n)).astype(np.int32)).cast(pa.date32())})


def supplier_table(sf: float, seed: int = 10) -> pa.Table:
    n = max(1, int(SUPPLIER_PER_SF * sf))
    rng = np.random.default_rng(seed)
    return pa.table({"s_suppkey": np.arange(1, n + 1, dtype=np.int64), "s_nationkey": rng.integers(0, 25, n).astype(np.int64)})


NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]   # TPC-H nation -> region


def nation_table() -> pa.Table:
    return pa.table({"n_nationkey": np.arange(25, dtype=np.int64), "n_name": np.arange(25, dtype=np.int32),
                     "n_regionkey": np.array(NATION_REGION, dtype=np.int64)})


def region_table() -> pa.Table:
    return pa.table({"r_regionkey": np.arange(5, dtype=np.int64), "r_name": np.arange(5, dtype=np.int32)})


# ------------------------------------------------------------------------------------------ Q3 / Q5 plans
def q3_plan(customer: SparkPlan, orders: SparkPlan, lineitem: SparkPlan, partial_final: bool = True) -> SparkPlan:
    """q3.sql as a physical plan.  Shape follows tpch-plan-stability/q3/simplified.txt (two hash joins feeding a
    Partial/Final aggregate and TakeOrderedAndProject); the build sides are the filtered customer and the
    customer-filtered orders (the golden plan's no-statistics choice of lineitem as a broadcast side is a
    planner decision, not a semantic one: inner joins commute)."""
    from .execution import BroadcastHashJoinExec, TakeOrderedAndProjectExec
    from .expressions import SortOrder
    cust = ProjectExec(["c_custkey"], FilterExec(col("c_mktsegment").eq(Literal(Q3_SEGMENT)), customer))
    ord_f = FilterExec(col("o_orderdate") < Literal(Q3_DATE), orders)
    j1 = BroadcastHashJoinExec(["o_custkey"], ["c_custkey"], "inner", "right", ord_f, cust)
    j1p = ProjectExec(["o_orderkey", "o_orderdate", "o_shippriority"], j1)
    li = ProjectExec(["l_orderkey", "l_extendedprice", "l_discount"], FilterExec(col("l_shipdate") > Literal(Q3_DATE), lineitem))
    j2 = BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "inner", "right", li, j1p)
    keys = ["l_orderkey", "o_orderdate", "o_shippriority"]
    aggs = [(Sum(col("l_extendedprice") * (Literal(1) - col("l_discount"))), "revenue")]
    if partial_final:
        agg = HashAggregateExec(keys, aggs, HashAggregateExec(keys, aggs, j2, mode="partial"), mode="final")
    else:
        agg = HashAggregateExec(keys, aggs, j2, mode="complete")
    return TakeOrderedAndProjectExec(10, [SortOrder("revenue", False), SortOrder("o_orderdate", True)],
                                     ["l_orderkey", "revenue", "o_orderdate", "o_shippriority"], agg)


def q5_plan(customer, orders, lineitem, supplier, nation, region, runtime_filter=True) -> SparkPlan:
    """q5.sql: six-way join, filter on region and order date, group by nation name, order by revenue desc.
    runtime_filter: the supplier side (selective: one region of five) also filters lineitem's l_suppkey below the orders join, the way
    InjectRuntimeFilter (sql/catalyst/.../optimizer/InjectRuntimeFilter.scala) puts a might-contain filter on the application side
    of a join whose creation side has a selective predicate; the supplier subplan is a ReusedExchange of its two consumers."""
    from .execution import BroadcastHashJoinExec, ReusedExchangeExec, RuntimeFilter
    reg = ProjectExec(["r_regionkey"], FilterExec(col("r_name").eq(Literal(Q5_REGION)), region))
    nat = ProjectExec(["n_nationkey", "n_name"], BroadcastHashJoinExec(["n_regionkey"], ["r_regionkey"], "inner", "right", nation, reg))
    if runtime_filter:
        nat = ReusedExchangeExec(nat, uses=2)
    sup = ProjectExec(["s_suppkey", "s_nationkey", "n_name"],
                      BroadcastHashJoinExec(["s_nationkey"], ["n_nationkey"], "inner", "right", supplier, nat))
    rfs = None
    cust = ProjectExec(["c_custkey", "c_nationkey"], customer)
    if runtime_filter:
        sup = ReusedExchangeExec(sup, uses=2)
        rfs = [RuntimeFilter("l_suppkey", "s_suppkey", sup)]
        # c_nationkey = s_nationkey with s_nationkey confined to the region's nations: the same rule's IN-subquery on the customer
        # side, planned as the semi join it logically is (the relation on 25 dense keys is an exact bitmap)
        cust = BroadcastHashJoinExec(["c_nationkey"], ["n_nationkey"], "left_semi", "right", cust, ProjectExec(["n_nationkey"], nat))
    ord_f = FilterExec((col("o_orderdate") >= Literal(Q5_DATE_LO)) & (col("o_orderdate") < Literal(Q5_DATE_HI)), orders)
    oc = ProjectExec(["o_orderkey", "c_nationkey"], BroadcastHashJoinExec(["o_custkey"], ["c_custkey"], "inner", "right", ord_f, cust))
    lo = ProjectExec(["l_suppkey", "c_nationkey", "l_extendedprice", "l_discount"],
                     BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "inner", "right",
                                           ProjectExec(["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"], lineitem), oc,
                                           runtimeFilters=rfs))
    # q5.sql: l_suppkey = s_suppkey AND c_nationkey = s_nationkey -- both are equi-join keys of the supplier join
    # (ExtractEquiJoinKeys, sql/catalyst/.../planning/patterns.scala); together they need more than 64 bits: the wide-key relation
    same_nation = BroadcastHashJoinExec(["l_suppkey", "c_nationkey"], ["s_suppkey", "s_nationkey"], "inner", "right", lo, sup)
    aggs = [(Sum(col("l_extendedprice") * (Literal(1) - col("l_discount"))), "revenue")]
    agg = HashAggregateExec(["n_name"], aggs, HashAggregateExec(["n_name"], aggs, same_nation, mode="partial"), mode="final")
    return SortExec([("revenue", False, False)], agg)


# ------------------------------------------------------------------------------------------ synthetic dataset (include/sb_synth.h)
SYNTH_TABLE = {"lineitem": 1, "orders": 2, "customer": 3, "supplier": 4}
SYNTH_COLUMNS = {
    "lineitem": ["l_orderkey", "l_partkey", "l_suppkey", "l_linenumber", "l_quantity", "l_extendedprice", "l_discount", "l_tax",
                 "l_returnflag", "l_linestatus", "l_shipdate", "l_commitdate", "l_receiptdate"],
    "orders": ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"],
    "customer": ["c_custkey", "c_mktsegment", "c_nationkey"],
    "supplier": ["s_suppkey", "s_nationkey"],
}
_SYNTH_NP = {8: np.int64, 4: np.int32, 1: np.int8}
_F64_COLS = {"l_quantity", "l_extendedprice", "l_discount", "l_tax"}
_DATE_COLS = {"l_shipdate", "l_commitdate", "l_receiptdate", "o_orderdate"}
Q1_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
CONFIG4_COLUMNS = SYNTH_COLUMNS["lineitem"]          # the 74 B/row fixed-width lineitem row of BASELINE.json configs[3]
CONFIG4_BYTES_PER_ROW = 74


def synth_width(name):
    if name in _F64_COLS or name.endswith("key"):
        return 8
    if name in ("l_returnflag", "l_linestatus", "c_mktsegment"):
        return 1
    return 4


def synth_dtype(name):
    return np.float64 if name in _F64_COLS else _SYNTH_NP[synth_width(name)]


def synth_rows(table: str, n_orders: int) -> int:
    if table == "lineitem":
        return (n_orders // 7) * 28 + [0, 1, 3, 6, 10, 15, 21][n_orders % 7]
    if table == "orders":
        return n_orders
    return max(1, n_orders // (10 if table == "customer" else 150))


def synth_batch(table: str, columns, n_orders: int, seed: int = 42, first_row: int = 0, nrows: int = None, stream=None):
    """Columns of the synthetic dataset generated on the GPU (sb_synth_table): a ColumnarBatch resident in HBM."""
    import ctypes as C
    from . import _capi as capi
    from .columnar import ColumnarBatch, _h
    lib = capi.init()
    ids = [SYNTH_COLUMNS[table].index(c) for c in columns]
    arr = (C.c_int32 * len(ids))(*ids)
    if nrows is None:
        nrows = synth_rows(table, n_orders) - first_row
    h = C.c_void_p()
    capi.check(lib.sb_synth_table(SYNTH_TABLE[table], arr, len(ids), n_orders, first_row, nrows, seed, _h(stream), C.byref(h)))
    ats = [pa.date32() if c in _DATE_COLS else pa.from_numpy_dtype(synth_dtype(c)) for c in columns]
    return ColumnarBatch(h, list(columns), ats)
