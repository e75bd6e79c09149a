#!/bin/bash
# 4 x B200: multi-GPU parity tests (world 2 and 4 cases) + a short N=4 bench (weak)
TAG=${1:-m4}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "4-" 2>&1 | tail -6) | tee gpurun_out/${TAG}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 30 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("N=4 value", round(d["value"], 1), "e2e", round(e["value"], 1), round(e["ms_per_step"], 3), "ms chunks", e.get("pipelined_chunks"), "serial", e["serial_transfers"]["value"], "full", e["full_signature"]["value"], "parity", d.get("multi_gpu_parity"))
PY
tail -3 gpurun_out/${TAG}_bench.err
