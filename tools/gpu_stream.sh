#!/bin/bash
# B200 check of the task-stream kernel (C3D20): bitwise parity against the two-phase path, then bench lines for a few settings
TAG=${1:-st}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/${TAG}_tests.log
run() {
  env "$@" timeout 300 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 10 --no-cpu --no-e2e --no-extra 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$*', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
" | tee -a gpurun_out/${TAG}_bench.log
}
run EWB_STREAM=1
run EWB_STREAM=1 EWB_STREAM_DISCARD=0
run EWB_STREAM=1 EWB_ELEMENT_ORDER=none
run EWB_STREAM=0
tail -5 gpurun_out/${TAG}_err.log
