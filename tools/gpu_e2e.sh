#!/bin/bash
# B200: host-call tests + the e2e figures of the default workload
TAG=${1:-e2e}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -x -q -k "chunk_pipelined or compute_host or plugin or cuda_backend or job" 2>&1 | tail -6) | tee gpurun_out/${TAG}_tests.log
env $EXTRA_ENV timeout 600 python bench.py --no-extra --no-cpu --steps 30 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), round(e["ms_per_step"], 3), "ms chunks", e.get("pipelined_chunks"))
print("serial", e["serial_transfers"], "\nfull", e["full_signature"]["value"], "pageable", e["pageable_numpy_inputs"]["value"])
PY
tail -3 gpurun_out/${TAG}_bench.err
