#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r15_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r15_tests.log
for c in strings window decimal join; do
python tools/op_bench.py $c > gpurun_out/r15_op_$c.jsonl 2> gpurun_out/r15_op_$c.err; echo "$c rc=$?"; cut -c1-520 gpurun_out/r15_op_$c.jsonl; tail -2 gpurun_out/r15_op_$c.err
done
python __graft_entry__.py smoke 2>&1 | tail -2
