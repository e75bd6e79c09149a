#!/bin/bash
# B200: C3D20 "row" two-phase path (EWB_C3D20_ROWS=1) against the half-block two-phase path: parity, bench, DRAM bytes
TAG=${1:-rows}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_stream.py -m gpu -x -q -k "env5 or flags or oracle" 2>&1 | tail -3) | tee gpurun_out/${TAG}_tests.log
for cfg in EWB_C3D20_ROWS=1 EWB_C3D20_ROWS=0; do
  env $cfg timeout 200 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 10 --no-cpu --no-e2e --no-extra 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$cfg', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms launches', d['gpu_launches'])
" | tee -a gpurun_out/${TAG}_bench.log
done
EWB_C3D20_ROWS=1 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"rowElementsKernel|rowGatherKernel" -c 2 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra 2>&1 | grep -E "rowElements|rowGather|dram__|gpu__time" | cut -c1-120 | tee gpurun_out/${TAG}_ncu.log
