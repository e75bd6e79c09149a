#!/bin/bash
# B200: chunk-by-chunk launches with their own x-chunk count (EWB_PIPE_CHUNKS): bitwise against the single launch? e2e?
TAG=${1:-pc}
mkdir -p gpurun_out
EWB_PIPE_CHUNKS=3 EWB_HOST_REGISTER_MIN=0 python - <<'PY' 2>&1 | tail -4 | tee gpurun_out/${TAG}_bitwise.log
import numpy as np, torch
from edelweissfe_b200 import ElementAssembly, box_mesh
for material, props in (("linearelastic", [2.1e4, 0.22]), ("vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400])):
    n = (25, 9, 17)
    coords, conn = box_mesh(*n)
    rng = np.random.default_rng(8)
    a0 = ElementAssembly("C3D8", conn, coords, material, props, box=n)
    a1 = ElementAssembly("C3D8", conn, coords, material, props, box=n)
    U_n = 1e-3 * rng.standard_normal(a0.nDof); dU = 3e-3 * rng.standard_normal(a0.nDof)
    a0.begin_increment(U_n); a1.begin_increment(U_n)
    P0, f0 = a0.compute_host_increment(dU); P0 = P0.copy()
    P1, f1 = a1.compute_host_increment_pipelined(dU)
    print(material, "chunks", a1.x_chunks(), "K equal", bool(torch.equal(a0.csr_data, a1.csr_data)), "P equal", np.array_equal(P0, P1), "F equal", bool(torch.equal(a0.F, a1.F)),
          "max rel K diff", float((a0.csr_data - a1.csr_data).abs().max() / a0.csr_data.abs().max()))
PY
for c in 0 6; do
  EWB_PIPE_CHUNKS=$c timeout 200 python bench.py --no-extra --no-cpu --steps 30 2>gpurun_out/${TAG}_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('EWB_PIPE_CHUNKS=$c value', round(d['value'], 1), 'e2e', round(e['value'], 1), round(e['ms_per_step'], 3), 'ms chunks', e.get('pipelined_chunks'))
" | tee -a gpurun_out/${TAG}.log
done
