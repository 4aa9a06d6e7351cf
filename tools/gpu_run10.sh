#!/bin/bash
# 2-GPU validation: fused / two-step exchange vs the oracle, then the bench at N=2 (all legs incl. shuffle)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r10_mgc.log 2>&1
echo "multi_gpu_check rc=$?"; tail -15 gpurun_out/r10_mgc.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r10_bench_n2.json 2> gpurun_out/r10_bench_n2.err
echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/r10_bench_n2.json; tail -8 gpurun_out/r10_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r10_ref_n2.json 2> gpurun_out/r10_ref_n2.err
echo "ref n2 rc=$?"; tail -c 1500 gpurun_out/r10_ref_n2.json; tail -3 gpurun_out/r10_ref_n2.err
