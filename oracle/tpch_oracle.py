"""CPU ORACLE (test infrastructure): TPC-H Q1/Q3/Q5 composed from the operator oracles.

TPC-H answers are "parity unpinned" in the reference (it ships the queries, sql/core/src/test/resources/tpch/,
but no data and no expected results), so these pipelines are the operator-by-operator restatement of the golden
plan shapes (tpch-plan-stability/q{1,3,5}/simplified.txt), each operator pinned separately.
Constants are passed in by the caller (the tests import them from spark_b200.tpch, the shared query definition).
"""
import numpy as np

from . import oracle as O

_REV = ("mul", ("col", "l_extendedprice"), ("sub", ("lit", 1.0), ("col", "l_discount")))


def q1(lineitem, cutoff, sort=True):
    f = O.filter_table(lineitem, ("le", ("col", "l_shipdate"), ("lit", cutoff, np.int32)))
    p = O.project(f, [(c, ("col", c)) for c in ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount"]] +
                  [("disc_price", _REV), ("charge", ("mul", _REV, ("add", ("lit", 1.0), ("col", "l_tax"))))])
    aggs = [("sum", "l_quantity", "sum_qty"), ("sum", "l_extendedprice", "sum_base_price"), ("sum", "disc_price", "sum_disc_price"),
            ("sum", "charge", "sum_charge"), ("avg", "l_quantity", "avg_qty"), ("avg", "l_extendedprice", "avg_price"),
            ("avg", "l_discount", "avg_disc"), ("count_star", None, "count_order")]
    out = O.hash_aggregate(p, ["l_returnflag", "l_linestatus"], aggs)
    return O.sort(out, [("l_returnflag", True, True), ("l_linestatus", True, True)]) if sort else out


def q3(customer, orders, lineitem, segment, date, limit=10):
    cust = O.project(O.filter_table(customer, ("eq", ("col", "c_mktsegment"), ("lit", segment, np.int32))), [("c_custkey", ("col", "c_custkey"))])
    ord_f = O.filter_table(orders, ("lt", ("col", "o_orderdate"), ("lit", date, np.int32)))
    j1 = O.hash_join(ord_f, cust, ["o_custkey"], ["c_custkey"], "inner").select(["o_orderkey", "o_orderdate", "o_shippriority"])
    li = O.filter_table(lineitem, ("gt", ("col", "l_shipdate"), ("lit", date, np.int32))).select(["l_orderkey", "l_extendedprice", "l_discount"])
    j2 = O.hash_join(li, j1, ["l_orderkey"], ["o_orderkey"], "inner")
    p = O.project(j2, [("l_orderkey", ("col", "l_orderkey")), ("o_orderdate", ("col", "o_orderdate")),
                       ("o_shippriority", ("col", "o_shippriority")), ("rev", _REV)])
    agg = O.hash_aggregate(p, ["l_orderkey", "o_orderdate", "o_shippriority"], [("sum", "rev", "revenue")])
    top = O.take_ordered(agg, [("revenue", False, False), ("o_orderdate", True, True)], limit)
    return top.select(["l_orderkey", "revenue", "o_orderdate", "o_shippriority"]), agg


def q5(customer, orders, lineitem, supplier, nation, region, region_code, date_lo, date_hi):
    reg = O.filter_table(region, ("eq", ("col", "r_name"), ("lit", region_code, np.int32))).select(["r_regionkey"])
    nat = O.hash_join(nation, reg, ["n_regionkey"], ["r_regionkey"], "inner").select(["n_nationkey", "n_name"])
    sup = O.hash_join(supplier, nat, ["s_nationkey"], ["n_nationkey"], "inner").select(["s_suppkey", "s_nationkey", "n_name"])
    ord_f = O.filter_table(orders, ("and", ("ge", ("col", "o_orderdate"), ("lit", date_lo, np.int32)),
                                    ("lt", ("col", "o_orderdate"), ("lit", date_hi, np.int32))))
    oc = O.hash_join(ord_f, customer.select(["c_custkey", "c_nationkey"]), ["o_custkey"], ["c_custkey"], "inner").select(["o_orderkey", "c_nationkey"])
    lo = O.hash_join(lineitem.select(["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"]), oc, ["l_orderkey"], ["o_orderkey"],
                     "inner").select(["l_suppkey", "c_nationkey", "l_extendedprice", "l_discount"])
    los = O.hash_join(lo, sup, ["l_suppkey"], ["s_suppkey"], "inner")
    same = O.filter_table(los, ("eq", ("col", "c_nationkey"), ("col", "s_nationkey")))
    p = O.project(same, [("n_name", ("col", "n_name")), ("rev", _REV)])
    agg = O.hash_aggregate(p, ["n_name"], [("sum", "rev", "revenue")])
    return O.sort(agg, [("revenue", False, False)])


# ------------------------------------------------------------------------------------------ synthetic dataset on the host
def synth_host(table: str, columns, n_orders: int, seed: int = 42, first_row: int = 0, nrows: int = None, out: dict = None) -> dict:
    """Columns of the synthetic dataset (include/sb_synth.h) filled on the host by so_synth_fill (OpenMP, parallel first
    touch): the CPU-side twin of spark_b200.tpch.synth_batch, bit-identical by construction and checked in the tests."""
    from spark_b200 import tpch as T        # names / ids of the synthetic columns only (no GPU code is touched)
    L = O.lib()
    if nrows is None:
        nrows = T.synth_rows(table, n_orders) - first_row
    res = {}
    for c in columns:
        a = out[c] if out is not None else np.empty(nrows, T.synth_dtype(c))
        L.so_synth_fill(T.SYNTH_TABLE[table], T.SYNTH_COLUMNS[table].index(c), n_orders, first_row, nrows, seed, a.ctypes.data)
        res[c] = a
    return res


def synth_arrow(table: str, columns, n_orders: int, seed: int = 42):
    """Host Arrow table of the synthetic dataset (input of the oracle pipelines in the tests)."""
    import pyarrow as pa
    from spark_b200 import tpch as T
    cols = synth_host(table, columns, n_orders, seed)
    return pa.table({c: (pa.array(cols[c]).cast(pa.date32()) if c in T._DATE_COLS else cols[c]) for c in columns})
