#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python tools/exp/sort_variants.py > gpurun_out/r4_sort_variants.jsonl 2> gpurun_out/r4_sort_variants.err; echo "variants rc=$?"; cat gpurun_out/r4_sort_variants.jsonl; tail -3 gpurun_out/r4_sort_variants.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 8 -c 2 -f -o gpurun_out/r4_onesweep python tools/exp/sort_variants.py 0 > gpurun_out/r4_ncu_onesweep.log 2>&1; echo "ncu onesweep rc=$?"; tail -3 gpurun_out/r4_ncu_onesweep.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_q3_sf10.csv python bench.py --sf 10 --legs q3 --steps 1 --warmup 1 --leg-steps 1 --e2e-steps 1 --no-cpu-baseline --no-verify > gpurun_out/r4_ncu_q3.log 2>&1; echo "ncu q3 rc=$?"; tail -2 gpurun_out/r4_ncu_q3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_update_kernel -s 6 -c 1 -f -o gpurun_out/r4_agg_update python bench.py --sf 10 --legs q1 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-verify > gpurun_out/r4_ncu_agg.log 2>&1; echo "ncu agg rc=$?"; tail -2 gpurun_out/r4_ncu_agg.log
ls -la gpurun_out/*.ncu-rep
