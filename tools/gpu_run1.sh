#!/bin/bash
# round-2 GPU session 1: new tests, then the new bench at SF10 and SF100
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/r1_tests_new.log 2>&1
echo "new tests rc=$?" 
tail -15 gpurun_out/r1_tests_new.log
timeout 600 python bench.py --sf 10 --steps 10 > gpurun_out/r1_bench_sf10.json 2> gpurun_out/r1_bench_sf10.err
echo "bench sf10 rc=$?"; tail -c 3000 gpurun_out/r1_bench_sf10.json; tail -5 gpurun_out/r1_bench_sf10.err
timeout 900 python bench.py --steps 10 > gpurun_out/r1_bench_sf100.json 2> gpurun_out/r1_bench_sf100.err
echo "bench sf100 rc=$?"; tail -c 4000 gpurun_out/r1_bench_sf100.json; tail -5 gpurun_out/r1_bench_sf100.err
python -m pytest tests -x -q -m gpu > gpurun_out/r1_tests_all.log 2>&1
echo "all tests rc=$?"; tail -8 gpurun_out/r1_tests_all.log
