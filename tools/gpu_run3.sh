#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sort.py tests/test_gpu_join.py tests/test_gpu_round2.py tests/test_gpu_scan.py -x -q -m gpu > gpurun_out/r3_tests_new.log 2>&1
echo "new tests rc=$?"; tail -30 gpurun_out/r3_tests_new.log
timeout 600 python tools/op_bench.py sort > gpurun_out/r3_op_sort.jsonl 2> gpurun_out/r3_op_sort.err; echo "sort rc=$?"; cat gpurun_out/r3_op_sort.jsonl; tail -3 gpurun_out/r3_op_sort.err
timeout 600 python tools/op_bench.py scan > gpurun_out/r3_op_scan.jsonl 2> gpurun_out/r3_op_scan.err; echo "scan rc=$?"; cat gpurun_out/r3_op_scan.jsonl; tail -3 gpurun_out/r3_op_scan.err
timeout 600 python tools/op_bench.py join > gpurun_out/r3_op_join.jsonl 2> gpurun_out/r3_op_join.err; echo "join rc=$?"; cat gpurun_out/r3_op_join.jsonl; tail -3 gpurun_out/r3_op_join.err
timeout 600 python tools/op_bench.py partition > gpurun_out/r3_op_part.jsonl 2> gpurun_out/r3_op_part.err; echo "part rc=$?"; cat gpurun_out/r3_op_part.jsonl; tail -3 gpurun_out/r3_op_part.err
python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests_all.log 2>&1
echo "all tests rc=$?"; tail -15 gpurun_out/r3_tests_all.log
