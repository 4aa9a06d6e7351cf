#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r18_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r18_tests.log
timeout 1500 python bench.py > gpurun_out/r18_bench_n1.json 2> gpurun_out/r18_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r18_bench_n1.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "", l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r18_bench_n1.err
python tools/op_bench.py join > gpurun_out/r18_op_join.jsonl 2> gpurun_out/r18_op_join.err; echo "join rc=$?"; cut -c1-520 gpurun_out/r18_op_join.jsonl
python __graft_entry__.py smoke 2>&1 | tail -1
