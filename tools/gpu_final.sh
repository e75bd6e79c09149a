#!/bin/bash
# B200, final build of the round: smoke(), the whole GPU test-suite, the default bench run, the ncu launch list of a short bench run
# and one --set full capture of the bench kernel
TAG=${1:-final}
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6) | tee gpurun_out/${TAG}_smoke.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", round(e["value"], 1), round(e["ms_per_step"], 3), "ms; launches", d["gpu_launches"])
print({k: round(v["value"], 1) for k, v in d.get("extra_workloads", {}).items()}, d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowPipeKernel -c 1 -o gpurun_out/${TAG}_le -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
