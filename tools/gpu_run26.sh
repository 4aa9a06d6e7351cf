#!/bin/bash
# 4-GPU validation of what the driver runs at round end: bench.py --gpus 4 (all legs incl. shuffle), short
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --leg-steps 3 --e2e-steps 1 > gpurun_out/r26_bench_n4.json 2> gpurun_out/r26_bench_n4.err
echo "bench n4 rc=$?"; python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r26_bench_n4.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config','roofline','e2e','cpu_baseline')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "")
        if k=="shuffle": print("   ", json.dumps({a:b for a,b in l.items() if a not in ("kernel_ms_per_step",)})[:1500])
except Exception as e: print("ERR",e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r26_bench_n4.err | tail -8 | cut -c1-300
