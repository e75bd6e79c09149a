#!/bin/bash
# B200: task-stream kernel, schedule knobs (chunk size / gather delay / elements per task)
TAG=${1:-sw}
mkdir -p gpurun_out
run() {
  env "$@" timeout 200 python bench.py --workload boxgen100x100x50_c3d20_linearelastic --steps 10 --no-cpu --no-e2e --no-extra 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$*', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
" | tee -a gpurun_out/${TAG}_bench.log
}
for cfg in "$@"; do
  [ "$cfg" = "$TAG" ] && continue
  run $(echo $cfg | tr ',' ' ')
done
tail -3 gpurun_out/${TAG}_err.log
