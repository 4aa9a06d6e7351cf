export SB_OPBENCH_GROUPS=1024,65536,1048576,16777216
for cfg in "A=0"; do
  echo "== $cfg"
  env $cfg timeout 150 python tools/op_bench.py agg 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'][25:], 'ms=%.3f'%d['ms'], 'kern=%.3f'%d['kernels_ms']['agg_update'], 'frac=%.3f'%d['operator_frac_of_measured_peak'])
"
done
