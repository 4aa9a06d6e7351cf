"""How a plan with several double accumulators behaves between the dictionary tiers' reach and the shared-memory tier."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from op_bench import timed
from spark_b200 import _capi as capi
from spark_b200.columnar import ColumnarBatch, Stream
from spark_b200.execution import HashAggregateExec, LocalTableScanExec
from spark_b200.expressions import Average, Count, Sum, col
lib = capi.init(0); stream = Stream()
n = 32_000_000
rng = np.random.default_rng(1)
x, y, z = rng.random(n), rng.random(n), rng.random(n)
for groups in [int(g) for g in os.environ.get("SB_EXP_GROUPS", "4,8,12,20,50,200,2000").split(",")]:
    b = ColumnarBatch.from_numpy({"k": rng.integers(0, groups, n), "x": x, "y": y, "z": z}, stream)
    stream.synchronize()
    agg = HashAggregateExec(["k"], [(Sum(col("x")), "sx"), (Sum(col("y")), "sy"), (Average(col("z")), "az"), (Count(), "c")], LocalTableScanExec(b))
    ms, k = timed(lib, stream, lambda: agg.executeColumnar(stream).close(), ["agg_update"])
    print(json.dumps({"groups": groups, "rows": n, "ms": round(ms, 3), "agg_update_ms": round(k["agg_update"][0], 3), "GBps": round(n * 32 / ms / 1e6, 1)}), flush=True)
    b.close()
