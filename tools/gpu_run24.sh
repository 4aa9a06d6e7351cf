#!/bin/bash
# onesweep after the instruction diet (tests, variants, op bench); Q5 launch list + ncu of the runtime-filter candidate pass
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_window.py -x -q -m gpu > gpurun_out/r24_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r24_tests.log
timeout 600 python tools/exp/sort_variants.py > gpurun_out/r24_variants.jsonl 2>&1; echo "variants rc=$?"; cut -c1-200 gpurun_out/r24_variants.jsonl
timeout 600 python tools/op_bench.py sort > gpurun_out/r24_op_sort.jsonl 2>&1; cut -c1-700 gpurun_out/r24_op_sort.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r24_launches_q5.csv python bench.py --legs q5 --steps 1 --warmup 3 --leg-steps 1 --no-cpu-baseline --no-verify --e2e-steps 1 > gpurun_out/r24_ncu_q5.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r24_launches_q5.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size')
out=[(r[ik][:90], r[ig], r[iv]) for r in rows[start+1:] if len(r)>iv]
# the last step: print the last 60 launches
for o in out[-70:]: print(o)
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:join_candidate_strided_kernel.*bool.1" -s 1 -c 1 -o gpurun_out/r24_cand_rf -f python bench.py --legs q5 --steps 1 --warmup 3 --leg-steps 1 --no-cpu-baseline --no-verify --e2e-steps 1 > gpurun_out/r24_ncu_rf.log 2>&1; echo "ncu rf rc=$?"
python tools/ncu_summary.py gpurun_out/r24_cand_rf.ncu-rep 599999994 > gpurun_out/r24_cand_rf_summary.txt 2>&1; cat gpurun_out/r24_cand_rf_summary.txt
ncu -i gpurun_out/r24_cand_rf.ncu-rep --page source --csv > gpurun_out/r24_cand_rf_source.csv 2>/dev/null
