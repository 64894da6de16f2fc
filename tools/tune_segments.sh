#!/bin/bash
# sweep of the segment schedule (first-segment rows x growth cap) on the bench workload
for s0 in 256 512 1024 4096; do for g in 8 16 32; do
  GVDB_SEG0_ROWS=$s0 GVDB_SEG_GROWTH=$g python bench.py --steps 10 --warmup 3 --stream-rows 0 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
st=d['roofline']['stage_ms_per_step']
print('seg0=$s0 growth=$g', 'qps=%.0f'%d['value'], 'ms/step=%.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k,v) for k,v in st.items()), 'launches', d['gpu_launches'], 'recall', d.get('recall_at_10'))"
done; done
