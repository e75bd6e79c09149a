#!/bin/bash
# A/B of tuning knobs of the LE sweep kernel on B200: bash tools/gpu_ab.sh TAG "ENV1=..;ENV2=.." "..." (one bench line per setting)
TAG=$1; shift
mkdir -p gpurun_out
for setting in "$@"; do
  (
    IFS=';' read -ra kvs <<< "$setting"
    for kv in "${kvs[@]}"; do [ -n "$kv" ] && export "$kv"; done
    t=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_shapes and sweep or full_size" 2>&1 | tail -1)
    for w in ${WORKLOADS:-boxgen100_c3d8_linearelastic}; do
    b=$(timeout 300 python bench.py --workload $w --steps 30 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'], round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
")
    echo "[$setting] $b | tests: $t" | tee -a gpurun_out/${TAG}_ab.log
    done
  )
done
