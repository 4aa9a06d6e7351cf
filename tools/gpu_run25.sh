#!/bin/bash
# runtime filters on the candidate list: join tests, Q3 / Q5, host profile of one Q5 step
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join.py tests/test_gpu_tpch.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r25_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r25_tests.log
timeout 900 python bench.py --legs q3,q5 --steps 3 --warmup 3 --leg-steps 5 --no-cpu-baseline --e2e-steps 1 --profile-host > gpurun_out/r25_bench_q35.json 2> gpurun_out/r25_bench_q35.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r25_bench_q35.json') if l.startswith('{')][-1]
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "", l.get("roofline",{}).get("frac"))
except Exception as e: print("ERR",e)
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r25_bench_q35.err | head -120 | cut -c1-180
