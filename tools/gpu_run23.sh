#!/bin/bash
# onesweep: every variant timed + checked against numpy's stable argsort, ncu of one pass of the first and second form;
# gather kernel (8 rows per thread): tests + Q3 / Q5
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_join.py tests/test_gpu_table.py tests/test_gpu_window.py -x -q -m gpu > gpurun_out/r23_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r23_tests.log
timeout 900 python tools/exp/sort_variants.py > gpurun_out/r23_variants.jsonl 2>&1; echo "variants rc=$?"; cat gpurun_out/r23_variants.jsonl | cut -c1-220
timeout 600 python tools/op_bench.py sort > gpurun_out/r23_op_sort.jsonl 2>&1; cut -c1-400 gpurun_out/r23_op_sort.jsonl
timeout 900 python bench.py --legs q3,q5 --steps 3 --warmup 3 --leg-steps 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r23_bench_q35.json 2> gpurun_out/r23_bench_q35.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r23_bench_q35.json') if l.startswith('{')][-1]
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "")
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r23_bench_q35.err
for V in 4 6; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 8 -c 1 -o gpurun_out/r23_onesweep_v$V -f python tools/exp/sort_variants.py $V ncu > gpurun_out/r23_ncu_v$V.log 2>&1; echo "ncu v$V rc=$?"
python tools/ncu_summary.py gpurun_out/r23_onesweep_v$V.ncu-rep 25000000 > gpurun_out/r23_onesweep_v${V}_summary.txt 2>&1; cat gpurun_out/r23_onesweep_v${V}_summary.txt
ncu -i gpurun_out/r23_onesweep_v$V.ncu-rep --page source --csv > gpurun_out/r23_onesweep_v${V}_source.csv 2>/dev/null
done
