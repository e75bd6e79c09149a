#!/bin/bash
# bench-only A/B (no tests): bash tools/gpu_ab_fast.sh TAG "ENV=.." ...
TAG=$1; shift
mkdir -p gpurun_out
for setting in "$@"; do
  (
    IFS=';' read -ra kvs <<< "$setting"
    for kv in "${kvs[@]}"; do [ -n "$kv" ] && export "$kv"; done
    for w in ${WORKLOADS:-boxgen100_c3d8_linearelastic}; do
    b=$(timeout 300 python bench.py --workload $w --steps 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'], round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
")
    echo "[$setting] $b" | tee -a gpurun_out/${TAG}_ab.log
    done
  )
done
