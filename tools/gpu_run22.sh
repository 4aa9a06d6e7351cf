#!/bin/bash
# 4-GPU validation: exchanges vs the oracle, then the bench at N=4 (all legs incl. shuffle), short
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r22_mgc.log 2>&1
echo "multi_gpu_check rc=$?"; tail -12 gpurun_out/r22_mgc.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --leg-steps 3 --e2e-steps 1 > gpurun_out/r22_bench_n4.json 2> gpurun_out/r22_bench_n4.err
echo "bench n4 rc=$?"; tail -c 2500 gpurun_out/r22_bench_n4.json; tail -8 gpurun_out/r22_bench_n4.err
