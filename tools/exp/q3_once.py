import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spark_b200 import _capi as capi, tpch
from spark_b200.columnar import ColumnarBatch, Stream
from spark_b200.execution import LocalTableScanExec
lib = capi.init(0); stream = Stream()
sf = 10.0
customer, orders = tpch.customer_table(sf), tpch.orders_table(sf)
lineitem = tpch.lineitem_join_table(orders, sf)
scan = lambda t: LocalTableScanExec(ColumnarBatch.from_arrow(t, stream))
which = sys.argv[1] if len(sys.argv) > 1 else "q3"
if which == "q3":
    plan = tpch.q3_plan(scan(customer), scan(orders), scan(lineitem), True)
else:
    plan = tpch.q5_plan(scan(customer), scan(orders), scan(lineitem), scan(tpch.supplier_table(sf)), scan(tpch.nation_table()), scan(tpch.region_table()))
for _ in range(2): plan.executeColumnar(stream).close()
stream.synchronize()
