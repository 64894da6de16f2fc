#!/bin/bash
# one-box sweep of the scan's item shape (GVDB_TC_QB: query blocks per work item) and the sample fraction
# (GVDB_SAMPLE_DIV) on the bench workload (configs[1]); prints one line per setting
run() {
  env "$@" python bench.py --steps 20 --warmup 5 --stream-rows 0 --no-cpu --north-star 0 --recall-queries 0 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
st=d['roofline']['stage_ms_per_step']
print('$*', 'qps=%.0f'%d['value'], 'ms/step=%.4f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], ' '.join('%s=%.4f'%(k,v) for k,v in st.items() if v), 'reruns', d['roofline']['optimistic_reruns'])"
}
run GVDB_NOP=1
run GVDB_TC_QB=4
run GVDB_TC_QB=3
run GVDB_TC_QB=2
run GVDB_SAMPLE_DIV=8
run GVDB_SAMPLE_DIV=32
run GVDB_NOP=2
