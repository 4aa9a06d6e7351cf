#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_strings.py tests/test_gpu_window.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r17_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r17_tests.log
for v in 1 2 3; do
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline --e2e-steps 1 --config join_cand=$v > gpurun_out/r17_bench_q35_cand$v.json 2> gpurun_out/r17_bench_q35_cand$v.err
echo "bench cand=$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r17_bench_q35_cand$v.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
done
python tools/op_bench.py strings > gpurun_out/r17_op_strings.jsonl 2> gpurun_out/r17_op_strings.err; echo "strings rc=$?"; cut -c1-520 gpurun_out/r17_op_strings.jsonl; tail -2 gpurun_out/r17_op_strings.err
