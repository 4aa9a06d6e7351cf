for cfg in "SB_RG_GRID=37" "SB_RG_GRID=74" "SB_RG_GRID=111" "SB_RG_GRID=148" "SB_RG_GRID=222" "SB_RG_GRID=296"; do
  echo "== $cfg"
  env $cfg timeout 120 python tools/op_bench.py partition 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['case'][:22], 'ms=%.3f'%d['ms'], 'scatter=%.3f'%d['kernels_ms']['partition_scatter'], 'frac=%.3f'%d['operator_frac_of_measured_peak'])
    except Exception as e: print(l[:200])
"
done
