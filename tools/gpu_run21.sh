#!/bin/bash
# N-GPU validation (N = $1): exchanges vs the oracle, then Q1 + the shuffle leg
N=${1:-4}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r21_mgc_n$N.log 2>&1
echo "multi_gpu_check rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r21_mgc_n$N.log | tail -6 | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --legs q1,q3,shuffle > gpurun_out/r21_bench_n$N.json 2> gpurun_out/r21_bench_n$N.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r21_bench_n$N.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config','roofline','e2e','cpu_baseline')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "")
        if k=="shuffle": print("   ", l.get("nvlink"), {v:(x["ms_per_step"]) for v,x in l.get("variants",{}).items()})
except Exception as e: print("ERR",e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r21_bench_n$N.err | tail -5 | cut -c1-300
