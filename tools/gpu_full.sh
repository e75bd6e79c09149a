#!/bin/bash
# B200: the whole GPU test-suite, then the default bench run
TAG=${1:-full}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) | tee gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")})
print("e2e full", d["e2e"].get("full_signature", {}).get("value"), "pageable", d["e2e"]["pageable_numpy_inputs"]["value"])
print({k: round(v["value"], 1) for k, v in d.get("extra_workloads", {}).items()})
PY
tail -3 gpurun_out/${TAG}_bench.err
