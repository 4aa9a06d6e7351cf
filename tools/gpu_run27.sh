#!/bin/bash
# 4 GPUs after the per-rank CPU partition: Q1 + Q3 + shuffle, verification on (it is what loads the OpenMP runtime)
cd /root/repo
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os; print(len(os.sched_getaffinity(0)))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --leg-steps 3 --e2e-steps 1 --legs q1,q3,shuffle --no-cpu-baseline > gpurun_out/r27_bench_n4.json 2> gpurun_out/r27_bench_n4.err
echo "bench n4 rc=$?"; python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r27_bench_n4.json') if l.startswith('{')][-1]
    print({k:v for k,v in d.items() if k not in ('legs','config','roofline','e2e','cpu_baseline')})
    for k,l in d["legs"].items():
        print(k, round(l["ms_per_step"],3), l.get("step_ms"), l["verified"], {a:round(b,3) for a,b in l["kernel_ms_per_step"].items()} if "kernel_ms_per_step" in l else "")
        if k=="shuffle": print("   ", json.dumps(l.get("nvlink")), json.dumps(l.get("variants"))[:600])
except Exception as e: print("ERR",e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r27_bench_n4.err | tail -5 | cut -c1-300
