#!/usr/bin/env python
"""TPC-H Q1 / Q3 / Q5 physical plans end to end on one GPU, inputs resident in HBM (BASELINE.json `metric`: rows/s :=
lineitem rows / device time of the pipeline from HBM-resident columns to the final result).  One JSON line per query:
    python tools/tpch_bench.py [SF=10] > gpurun_out/tpch_sf10.jsonl
Timing: CUDA events on the plan's stream, best of 5 after 2 warm-ups, plus the per-kernel-family breakdown the runtime
keeps (sb_profile_get).  Beside each GPU line the same query through the CPU oracle pipeline (oracle/tpch_oracle.py:
numpy/pyarrow operators + the C restatement) on a bounded sample (SF capped at 1), host cores stated."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from op_bench import timed                                  # noqa: E402
from spark_b200 import _capi as capi                        # noqa: E402
from spark_b200 import tpch                                 # noqa: E402
from spark_b200.columnar import ColumnarBatch, Stream       # noqa: E402
from spark_b200.execution import LocalTableScanExec         # noqa: E402

KERNELS = ["agg_update", "agg_update_final", "join_build", "join_probe", "join_fill", "partition_scatter", "filter_project", "gather"]


def main():
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    cpu_sf = min(sf, 1.0)
    lib = capi.init(0)
    stream = Stream()

    def scan(t):
        return LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))

    def report(query, n_lineitem, plan, cpu_fn, cpu_rows):
        ms, k = timed(lib, stream, lambda: plan.executeColumnar(stream).close(), KERNELS)
        t0 = time.perf_counter()
        cpu_fn()
        cpu_s = time.perf_counter() - t0
        print(json.dumps({"query": query, "sf": sf, "lineitem_rows": n_lineitem, "ms": ms, "rows_per_s": n_lineitem / (ms / 1e3),
                          "kernels_ms": {a: round(b[0], 4) for a, b in k.items() if b[1]},
                          "kernel_launches": {a: b[1] for a, b in k.items() if b[1]},
                          "cpu_oracle": {"sf": cpu_sf, "lineitem_rows": cpu_rows, "seconds": cpu_s, "rows_per_s": cpu_rows / cpu_s,
                                         "cores": len(os.sched_getaffinity(0)), "kind": "port (numpy/pyarrow operators + C restatement)"}}),
              flush=True)

    from oracle import tpch_oracle as TO
    # ---- Q1 ------------------------------------------------------------------------------------------------------------
    n1 = int(6_001_215 * sf)
    t1 = tpch.lineitem_q1_table(n1, seed=42)
    t1s = tpch.lineitem_q1_table(int(6_001_215 * cpu_sf), seed=42)
    plan = tpch.q1_final_plan(tpch.q1_partial_plan(scan(t1), fused=True), sort=True)
    report("q1", n1, plan, lambda: TO.q1(t1s, tpch.Q1_CUTOFF), t1s.num_rows)
    del plan, t1
    # ---- Q3 / Q5 tables ------------------------------------------------------------------------------------------------
    customer, orders = tpch.customer_table(sf), tpch.orders_table(sf)
    lineitem = tpch.lineitem_join_table(orders, sf)
    supplier, nation, region = tpch.supplier_table(sf), tpch.nation_table(), tpch.region_table()
    c_s, o_s = tpch.customer_table(cpu_sf), tpch.orders_table(cpu_sf)
    l_s = tpch.lineitem_join_table(o_s, cpu_sf)
    s_s = tpch.supplier_table(cpu_sf)
    sc, so, sl, ss, sn, sr = [scan(t) for t in (customer, orders, lineitem, supplier, nation, region)]
    report("q3", lineitem.num_rows, tpch.q3_plan(sc, so, sl, True),
           lambda: TO.q3(c_s, o_s, l_s, tpch.Q3_SEGMENT, tpch.Q3_DATE), l_s.num_rows)
    report("q5", lineitem.num_rows, tpch.q5_plan(sc, so, sl, ss, sn, sr),
           lambda: TO.q5(c_s, o_s, l_s, s_s, nation, region, tpch.Q5_REGION, tpch.Q5_DATE_LO, tpch.Q5_DATE_HI), l_s.num_rows)


if __name__ == "__main__":
    main()
