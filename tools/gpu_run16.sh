#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_window.py tests/test_gpu_join.py -x -q -m gpu > gpurun_out/r16_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r16_tests.log
for v in 0 1; do
timeout 900 python bench.py --steps 5 --legs q3,q5 --no-cpu-baseline --e2e-steps 1 --config join_cand=$v > gpurun_out/r16_bench_q35_cand$v.json 2> gpurun_out/r16_bench_q35_cand$v.err
echo "bench cand=$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r16_bench_q35_cand$v.json').read().strip().splitlines()[-1])
    for k,l in d["legs"].items(): print(k, l["ms_per_step"], l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()}, l["roofline"]["frac"])
except Exception as e: print("ERR",e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:join_candidate --csv --log-file gpurun_out/r16_cand_launches.csv python bench.py --legs q3 --steps 1 --warmup 3 --leg-steps 1 --no-cpu-baseline --no-verify --e2e-steps 1 --config join_cand=1 > /dev/null 2>&1; echo "ncu rc=$?"; grep "join_candidate" gpurun_out/r16_cand_launches.csv | tail -4 | cut -c1-300
