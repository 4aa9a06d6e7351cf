for v in 4p 4 8pf; do
  echo "== $v"
  SB_AGG_Q1_VARIANT=$v timeout 200 python bench.py --steps 100 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step=%.4f kernel_ms=%.4f frac=%.4f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
