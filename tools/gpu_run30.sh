#!/bin/bash
# the relation-ready event / no sync on return of the probe: targeted join tests (two streams, runtime filters, fixtures), TPC-H parity, Q3 / Q5
cd /root/repo
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_join.py tests/test_gpu_tpch.py -x -q -m gpu -k "another or runtime_filter or fixture or sparse_sorted or wider or tpch or q3 or q5 or empty" > gpurun_out/r30_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r30_tests.log
timeout 240 python bench.py --legs q3,q5 --steps 3 --warmup 3 --leg-steps 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r30_bench_q35.json 2> gpurun_out/r30_bench_q35.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r30_bench_q35.json') if l.startswith('{')][-1]
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "", l.get("roofline",{}).get("frac"))
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r30_bench_q35.err | cut -c1-300
