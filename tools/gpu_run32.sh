#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_sort.py -x -q -m gpu -k "rebuilt" > gpurun_out/r32_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r32_tests.log
