#!/usr/bin/env python
"""Onesweep tile-geometry experiment: sort 25 M full-range int64 keys with every variant (sb_config_set("sort_variant")).
With an argument N: run only variant N, twice (the shape an ncu capture wants)."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spark_b200 import _capi as capi
from spark_b200.columnar import ColumnarBatch, Stream
from spark_b200.execution import LocalTableScanExec, SortExec

lib = capi.init(0)
stream = Stream()
n = 25_000_000
rng = np.random.default_rng(0)
b = ColumnarBatch.from_numpy({"k": rng.integers(-2 ** 63, 2 ** 63 - 1, n)}, stream)
stream.synchronize()
srt = SortExec([("k", True, True)], LocalTableScanExec(b))
names = ["256x16", "512x8", "256x8", "512x16", "384x12", "1024x8"]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for v, name in enumerate(names):
    if only is not None and v != only:
        continue
    capi.config_set("sort_variant", v)
    reps = 2 if only is not None else 6
    best = None
    for i in range(reps):
        capi.check(lib.sb_profile_enable(1)); capi.check(lib.sb_profile_reset())
        srt.executeColumnar(stream).close()
        t, c = C.c_double(), C.c_int64()
        capi.check(lib.sb_profile_get(b"sort_passes", C.byref(t), C.byref(c)))
        if i > 0 and (best is None or t.value < best):
            best = t.value
    print(json.dumps({"variant": name, "sort_passes_ms": best, "pass_traffic_gbs": 8 * 24 * n / (best / 1e3) / 1e9}), flush=True)
