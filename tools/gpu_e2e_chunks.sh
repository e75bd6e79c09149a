#!/bin/bash
# B200: e2e of the default workload for several x-chunk counts of the fused kernel (EWB_CHUNKS, read at plan creation)
TAG=${1:-ch}
mkdir -p gpurun_out
for c in 0 4 6 8; do
  EWB_CHUNKS=$c timeout 300 python bench.py --no-extra --no-cpu --steps 30 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    e = d['e2e']
    print('EWB_CHUNKS=$c value', round(d['value'], 1), 'e2e', round(e['value'], 1), round(e['ms_per_step'], 3), 'ms chunks', e.get('pipelined_chunks'), 'pageable', round(e['pageable_numpy_inputs']['value'], 1))
" | tee -a gpurun_out/${TAG}.log
done
