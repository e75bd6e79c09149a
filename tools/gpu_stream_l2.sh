#!/bin/bash
# B200: does the row scratch stay in L2?  DRAM bytes + duration of the task-stream kernel (ncu, two metrics) for a few schedules
TAG=${1:-l2}
mkdir -p gpurun_out
for cfg in "$@"; do
  [ "$cfg" = "$TAG" ] && continue
  echo "== $cfg" | tee -a gpurun_out/${TAG}.log
  env $(echo $cfg | tr ',' ' ') timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:streamKernel -c 1 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra 2>&1 | grep -E "dram__|gpu__time" | tee -a gpurun_out/${TAG}.log
done
