#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_sort.py -x -q -m gpu -k "order_by_all" > gpurun_out/r34_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r34_tests.log
