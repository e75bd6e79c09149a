#!/bin/bash
# B200: C3D20 row two-phase path, gather variants; and the task-stream kernel with the same gather
TAG=${1:-rows2}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_stream.py -m gpu -x -q -k "env5 or env0" 2>&1 | tail -2) | tee gpurun_out/${TAG}_tests.log
for cfg in "EWB_C3D20_ROWS=1 EWB_ROWS_VARIANT=0" "EWB_C3D20_ROWS=1 EWB_ROWS_VARIANT=1" "EWB_C3D20_ROWS=1 EWB_ROWS_VARIANT=2" "EWB_STREAM=1"; do
  env $cfg timeout 200 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 10 --no-cpu --no-e2e --no-extra 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$cfg', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
" | tee -a gpurun_out/${TAG}_bench.log
done
