#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py tests/test_gpu_round2.py tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/r2_tests_new.log 2>&1
echo "new tests rc=$?"; tail -25 gpurun_out/r2_tests_new.log
timeout 900 python bench.py --steps 10 --legs q1 > gpurun_out/r2_bench_q1_sf100.json 2> gpurun_out/r2_bench_q1_sf100.err
echo "bench q1 sf100 rc=$?"; tail -c 5000 gpurun_out/r2_bench_q1_sf100.json; tail -5 gpurun_out/r2_bench_q1_sf100.err
timeout 300 python bench.py --steps 10 --legs q1 --sf 10 > gpurun_out/r2_bench_q1_sf10.json 2> gpurun_out/r2_bench_q1_sf10.err
echo "bench q1 sf10 rc=$?"; tail -c 2500 gpurun_out/r2_bench_q1_sf10.json; tail -5 gpurun_out/r2_bench_q1_sf10.err
