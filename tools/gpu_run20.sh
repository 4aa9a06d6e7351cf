#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_join.py tests/test_gpu_tpch.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r20_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r20_tests.log
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r20_bench_q35.json 2> gpurun_out/r20_bench_q35.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r20_bench_q35.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
