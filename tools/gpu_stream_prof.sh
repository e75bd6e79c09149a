#!/bin/bash
# B200: tests + bench + one ncu --set full capture of the task-stream kernel
TAG=${1:-sp}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -8) | tee gpurun_out/${TAG}_tests.log
run() {
  env "$@" timeout 200 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 10 --no-cpu --no-e2e --no-extra 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$*', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
" | tee -a gpurun_out/${TAG}_bench.log
}
run EWB_STREAM=1
run EWB_STREAM_DELAY=1000
run EWB_STREAM_NPT=16
timeout 400 ncu --set full --clock-control none --import-source on -k regex:streamKernel -c 1 -o gpurun_out/${TAG}_stream -f python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
