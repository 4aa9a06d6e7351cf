#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_join.py tests/test_gpu_tpch.py tests/test_gpu_sort.py tests/test_gpu_round2.py tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r6_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r6_tests.log
python tools/exp/sort_variants.py > gpurun_out/r6_sort_variants.jsonl 2> gpurun_out/r6_sort_variants.err; echo "variants rc=$?"; cat gpurun_out/r6_sort_variants.jsonl; tail -3 gpurun_out/r6_sort_variants.err
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline > gpurun_out/r6_bench_q3q5_sf100.json 2> gpurun_out/r6_bench_q3q5.err
echo "bench q3q5 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r6_bench_q3q5_sf100.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l["verified"], l["kernel_ms_per_step"], l["e2e"]["ms_per_step"], l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
tail -5 gpurun_out/r6_bench_q3q5.err
timeout 600 python tools/op_bench.py join > gpurun_out/r6_op_join.jsonl 2> gpurun_out/r6_op_join.err; echo "join rc=$?"; cat gpurun_out/r6_op_join.jsonl | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 8 -c 1 -f -o gpurun_out/r6_onesweep python tools/exp/sort_variants.py 0 > gpurun_out/r6_ncu_onesweep.log 2>&1; echo "ncu onesweep rc=$?"
