#!/bin/bash
# Quick B200 check of the sweep kernels: parity tests that exercise them + short device-resident bench lines.
# usage (through gpurun): bash tools/gpu_quick.sh TAG [extra env assignments]
TAG=${1:-quick}; shift
for kv in "$@"; do export "$kv"; done
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or full_size or golden" 2>&1 | tail -3) > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for w in boxgen100_c3d8_linearelastic boxgen200x100x100_c3d8_vonmises boxgen100_c3d8tl_neohookewa; do
  timeout 300 python bench.py --workload $w --steps 30 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'], round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms', 'frac', round(d['roofline']['frac'], 3))
" | tee -a gpurun_out/${TAG}_bench.log
done
