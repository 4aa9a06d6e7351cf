#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r7_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r7_tests.log
python tools/exp/sort_variants.py > gpurun_out/r7_sort_variants.jsonl 2> gpurun_out/r7_sort_variants.err; echo "variants rc=$?"; cat gpurun_out/r7_sort_variants.jsonl; tail -3 gpurun_out/r7_sort_variants.err
timeout 900 python bench.py --steps 5 --legs q3 --no-cpu-baseline > gpurun_out/r7_bench_q3_sf100.json 2> gpurun_out/r7_bench_q3.err
echo "bench q3 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r7_bench_q3_sf100.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["e2e"]["ms_per_step"], l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
tail -5 gpurun_out/r7_bench_q3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r7_launches_q3_sf100.csv python bench.py --legs q3 --steps 1 --warmup 1 --leg-steps 1 --e2e-steps 1 --no-cpu-baseline --no-verify > gpurun_out/r7_ncu_q3.log 2>&1; echo "ncu q3 rc=$?"; tail -2 gpurun_out/r7_ncu_q3.log | cut -c1-300
