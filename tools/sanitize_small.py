"""Small fused-kernel run for compute-sanitizer (memcheck / racecheck): both fused kernels, several x-chunks and ragged tiles.
usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edelweissfe_b200 import ElementAssembly, box_mesh  # noqa: E402

n = (9, 10, 9)
for elType, material, props, scale in (("C3D8", "linearelastic", [2.1e4, 0.22], 1e-3), ("C3D8", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], 5e-3)):
    coords, conn = box_mesh(*n, lX=9.0, lY=10.0, lZ=9.0, elType=elType)
    dU = scale * np.random.default_rng(0).standard_normal(3 * coords.shape[0])
    asm = ElementAssembly(elType, conn, coords, material, props, box=n)
    asm.U.copy_(torch.as_tensor(dU))
    asm.dU.copy_(torch.as_tensor(dU))
    for _ in range(2):
        asm.assemble()
    asm.poll()
    print(elType, material, "ok", float(asm.csr_data.abs().max()))
