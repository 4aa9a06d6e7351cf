#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_sort.py tests/test_gpu_window.py -x -q -m gpu -k "nulls_ordering or null_ordering_golden" > gpurun_out/r33_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r33_tests.log
