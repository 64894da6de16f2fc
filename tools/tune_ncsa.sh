python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0 8 9 10 11 12 13 14; do
  GVDB_SCAN_NCSA=$v python bench.py --steps 5 --warmup 2 --stream-rows 0 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ncsa=$v', 'qps=%.0f'%d['value'], 'ms/step=%.3f'%d['ms_per_step'], 'scan_ms=%.3f'%d['roofline']['stage_ms_per_step']['scan_ms'], 'recall', d.get('recall_at_10'))"
done
