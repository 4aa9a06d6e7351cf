#!/bin/bash
# ncu launch list (time only) of the default bench command, short
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r29_launches_bench.csv python bench.py --steps 2 --warmup 3 --leg-steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r29_ncu_bench.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r29_launches_bench.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[start+1:]:
    if len(r)>iv:
        k=r[ik].split('(')[0][:60]; tot[k]+=float(r[iv].replace(',','')); cnt[k]+=1
for k,v in tot.most_common(25): print("%10.3f ms %5d  %s"%(v/1e6,cnt[k],k))
PY
