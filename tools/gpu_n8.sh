#!/bin/bash
# 8 x B200: short weak-scaling bench of the default workload (device-resident value + e2e through the slab host calls)
TAG=${1:-n8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("N=8 value", round(d["value"], 1), "e2e", round(e["value"], 1), round(e["ms_per_step"], 3), "ms chunks", e.get("pipelined_chunks"), "serial", round(e["serial_transfers"]["value"], 1), "full", round(e["full_signature"]["value"], 1), "parity", d.get("multi_gpu_parity", {}).get("rel_err"))
PY
tail -2 gpurun_out/${TAG}_bench.err
