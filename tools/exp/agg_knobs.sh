export SB_OPBENCH_GROUPS=4,16,1024,4096,65536
for cfg in "A=0"; do
  echo "== $cfg"
  env $cfg timeout 150 python tools/op_bench.py agg 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'][25:], 'ms=%.3f'%d['ms'], 'kern=%.3f'%d['kernels_ms']['agg_update'], 'frac=%.3f'%d['operator_frac_of_measured_peak'])
"
done
