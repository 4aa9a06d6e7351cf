#!/usr/bin/env python
"""Onesweep tile-geometry experiment: sort 25 M full-range int64 keys with every variant (sb_config_set("sort_variant")).
With an argument N: run only variant N, twice (the shape an ncu capture wants).  Every variant is first checked against numpy's
stable argsort on a ragged size, with full-range keys and with heavily duplicated keys (tie order)."""
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
only = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else None
profile_mode = len(sys.argv) > 2 and sys.argv[2] == "ncu"

m = 3_000_017
checks = []
if not profile_mode:
    for label, keys in (("full", rng.integers(-2 ** 63, 2 ** 63 - 1, m)), ("dups", rng.integers(0, 70000, m)), ("tiny", rng.integers(0, 3, 5000))):
        cb = ColumnarBatch.from_numpy({"k": keys, "i": np.arange(len(keys), dtype=np.int32)}, stream)
        order = np.argsort(keys, kind="stable")
        checks.append((label, cb, SortExec([("k", True, True)], LocalTableScanExec(cb)), keys[order], order.astype(np.int32)))

for v, name in enumerate(names):
    if only is not None and v not in only:
        continue
    capi.config_set("sort_variant", v)
    ok = True
    for label, cb, s2, want_k, want_i in checks:
        out = s2.executeColumnar(stream)
        got_k = out.column_to_numpy(out.column_index("k"), stream)[0]
        got_i = out.column_to_numpy(out.column_index("i"), stream)[0]
        stream.synchronize()
        out.close()
        if not (np.array_equal(got_k, want_k) and np.array_equal(got_i, want_i)):
            ok = False
            print(json.dumps({"variant": name, "check": label, "equal": False}), flush=True)
    reps = 2 if profile_mode else 6
    best = None
    for i in range(reps):
        capi.check(lib.sb_profile_enable(1)); capi.check(lib.sb_profile_reset())
        srt.executeColumnar(stream).close()
        t, c = C.c_double(), C.c_int64()
        capi.check(lib.sb_profile_get(b"sort_passes", C.byref(t), C.byref(c)))
        if i > 0 and (best is None or t.value < best):
            best = t.value
    print(json.dumps({"variant": name, "v": v, "sort_passes_ms": best, "pass_traffic_gbs": 8 * 24 * n / (best / 1e3) / 1e9,
                      "checked": None if profile_mode else ok}), flush=True)
