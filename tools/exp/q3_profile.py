import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spark_b200 import _capi as capi, tpch
from spark_b200.columnar import ColumnarBatch, Stream
from spark_b200.execution import LocalTableScanExec
lib = capi.init(0); stream = Stream()
sf = 10.0
customer, orders = tpch.customer_table(sf), tpch.orders_table(sf)
lineitem = tpch.lineitem_join_table(orders, sf)
scan = lambda t: LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))
plan = tpch.q3_plan(scan(customer), scan(orders), scan(lineitem), True)
for _ in range(3): plan.executeColumnar(stream).close()
stream.synchronize()
t0 = time.perf_counter()
for _ in range(10): plan.executeColumnar(stream).close()
stream.synchronize()
print("wall ms per run", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): plan.executeColumnar(stream).close()
stream.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
