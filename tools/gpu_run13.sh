#!/bin/bash
# 2-GPU validation: exchanges (fused, two-step, strings) vs the oracle, then the bench at N=2 with every leg
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r13_mgc.log 2>&1
echo "multi_gpu_check rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r13_mgc.log | tail -12 | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r13_bench_n2.json 2> gpurun_out/r13_bench_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r13_bench_n2.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config','roofline','e2e','cpu_baseline')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "")
        if k=="shuffle": print("   ", l.get("nvlink"), {v:(x["ms_per_step"]) for v,x in l.get("variants",{}).items()})
except Exception as e: print("ERR",e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r13_bench_n2.err | tail -5 | cut -c1-300
