#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r12_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r12_tests.log
for v in 0 1; do
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline --e2e-steps 1 --config join_cand=$v > gpurun_out/r12_bench_q35_cand$v.json 2> gpurun_out/r12_bench_q35_cand$v.err
echo "bench cand=$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r12_bench_q35_cand$v.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r12_bench_q35_cand$v.err
done
python tools/op_bench.py join > gpurun_out/r12_op_join.jsonl 2> gpurun_out/r12_op_join.err; echo "join rc=$?"; cut -c1-620 gpurun_out/r12_op_join.jsonl
