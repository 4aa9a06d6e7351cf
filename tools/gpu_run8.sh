#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_join.py tests/test_gpu_round2.py tests/test_gpu_aggregate.py tests/test_gpu_sort.py -x -q -m gpu > gpurun_out/r8_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r8_tests.log
python tools/op_bench.py sort > gpurun_out/r8_op_sort.jsonl 2> gpurun_out/r8_op_sort.err; echo "sort rc=$?"; cut -c1-420 gpurun_out/r8_op_sort.jsonl
python tools/op_bench.py join > gpurun_out/r8_op_join.jsonl 2> gpurun_out/r8_op_join.err; echo "join rc=$?"; cut -c1-520 gpurun_out/r8_op_join.jsonl
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline > gpurun_out/r8_bench_q35_sf100.json 2> gpurun_out/r8_bench_q35.err
echo "bench q3q5 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r8_bench_q35_sf100.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["e2e"]["ms_per_step"], l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
tail -5 gpurun_out/r8_bench_q35.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_launches_q35_sf100.csv python bench.py --legs q3,q5 --steps 1 --warmup 1 --leg-steps 1 --e2e-steps 1 --no-cpu-baseline --no-verify > gpurun_out/r8_ncu_q35.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r8_ncu_q35.log | cut -c1-300
