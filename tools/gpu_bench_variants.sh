#!/bin/bash
# bench-only A/B of kernel variants: bash tools/gpu_bench_variants.sh TAG "VARIANT ..." [WORKLOADS...]
TAG=${1:-ab}; VARIANTS=$2; shift; shift
WL=${@:-boxgen100_c3d8_linearelastic}
mkdir -p gpurun_out
for v in $VARIANTS; do
  for w in $WL; do
    EWB_KERNEL=$v timeout 300 python bench.py --workload $w --steps 30 --no-cpu --no-e2e 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$v', d['config']['workload'], round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms', 'frac', round(d['roofline']['frac'], 3))
" | tee -a gpurun_out/${TAG}_bench.log
    tail -2 gpurun_out/${TAG}_err.log
  done
done
