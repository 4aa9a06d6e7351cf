#!/bin/bash
# exactly 256 buckets in one multisplit pass (the local split of an 8-rank exchange of 2048 partitions): partition tests + fused exchange
cd /root/repo
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_partition.py tests/test_gpu_round2.py -x -q -m gpu -k "256" > gpurun_out/r31_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r31_tests.log
