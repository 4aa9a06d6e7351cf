#!/bin/bash
# final validation on one GPU: the whole GPU suite, smoke, the default bench line, the reference arm
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r28_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r28_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r28_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r28_smoke.log
timeout 1200 python bench.py > gpurun_out/r28_bench_n1.json 2> gpurun_out/r28_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r28_bench_n1.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "", l.get("cpu_baseline",{}).get("value"), l.get("roofline",{}).get("frac"), l.get("e2e",{}).get("value"))
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r28_bench_n1.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r28_bench_ref.json 2> gpurun_out/r28_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r28_bench_ref.json
