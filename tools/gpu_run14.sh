#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_join.py tests/test_gpu_window.py tests/test_gpu_decimal.py tests/test_gpu_round2.py tests/test_gpu_tpch.py -x -q -m gpu > gpurun_out/r14_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r14_tests.log
# the default bench line (what the driver runs) and the reference arm
timeout 1500 python bench.py > gpurun_out/r14_bench_n1.json 2> gpurun_out/r14_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r14_bench_n1.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "", l.get("cpu_baseline"))
        if "variants" in l: print("   variants", {v:{a:(round(b,3) if isinstance(b,float) else b) for a,b in x.items()} for v,x in l["variants"].items()})
except Exception as e: print("ERR",e)
PY
tail -3 gpurun_out/r14_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r14_bench_ref.json 2> gpurun_out/r14_bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/r14_bench_ref.json
# ncu: DRAM traffic of the dominant kernel at the benchmarked size, and what limits the join candidate pass
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_update_kernel -s 4 -c 1 -o gpurun_out/r14_agg_sf100 -f python bench.py --legs q1 --steps 3 --warmup 3 --no-cpu-baseline --no-verify --e2e-steps 1 > gpurun_out/r14_ncu_agg.log 2>&1; echo "ncu agg rc=$?"
ncu -i gpurun_out/r14_agg_sf100.ncu-rep --page raw --csv > gpurun_out/r14_agg_sf100_raw.csv 2>/dev/null; head -c 300 gpurun_out/r14_agg_sf100_raw.csv | head -2 | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:join_candidate_kernel -s 3 -c 1 -o gpurun_out/r14_cand0 -f python bench.py --legs q3 --steps 1 --warmup 3 --leg-steps 1 --no-cpu-baseline --no-verify --e2e-steps 1 > gpurun_out/r14_ncu_cand0.log 2>&1; echo "ncu cand0 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:join_candidate_strided -s 3 -c 1 -o gpurun_out/r14_cand1 -f python bench.py --legs q3 --steps 1 --warmup 3 --leg-steps 1 --no-cpu-baseline --no-verify --e2e-steps 1 --config join_cand=1 > gpurun_out/r14_ncu_cand1.log 2>&1; echo "ncu cand1 rc=$?"
for f in cand0 cand1; do ncu -i gpurun_out/r14_$f.ncu-rep --page raw --csv > gpurun_out/r14_${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r14_*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
