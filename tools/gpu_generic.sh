#!/bin/bash
# B200 check of the arbitrary-mesh (generic) path: parity tests + C3D20 and forced-generic C3D8 bench lines
TAG=${1:-gen}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/${TAG}_tests.log
for args in "--workload boxgen100x100x50_c3d20_linearelastic" "--workload boxgen100_c3d8_linearelastic --generic"; do
  timeout 300 python bench.py $args --steps 10 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'], d['config'].get('path'), round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
" | tee -a gpurun_out/${TAG}_bench.log
done
