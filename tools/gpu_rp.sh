#!/bin/bash
# B200 check of the fused BoxGen kernels: parity tests + short device-resident bench lines per kernel variant.
# usage (through gpurun): bash tools/gpu_rp.sh TAG "VARIANT ..." [WORKLOADS...]   (VARIANT = v1 | rp484 | rp444 ...)
TAG=${1:-rp}; VARIANTS=${2:-"rp484 v1"}; shift; shift
WL=${@:-boxgen100_c3d8_linearelastic}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or golden" 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for v in $VARIANTS; do
  for w in $WL; do
    EWB_KERNEL=$v timeout 300 python bench.py --workload $w --steps 30 --no-cpu --no-e2e 2>gpurun_out/${TAG}_${v}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$v', d['config']['workload'], round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms', 'frac', round(d['roofline']['frac'], 3))
" | tee -a gpurun_out/${TAG}_bench.log
    tail -2 gpurun_out/${TAG}_${v}_err.log
  done
done
