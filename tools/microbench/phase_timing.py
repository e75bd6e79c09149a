"""Per-warp phase timing of the sweep kernel (debug build with -DEWB_TIMING, tools/microbench/libewb_timing.so)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from edelweissfe_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tools", "microbench", "libewb_timing.so")
import torch  # noqa: E402

from edelweissfe_b200 import ElementAssembly, box_mesh  # noqa: E402

n = (100, 100, 100)
coords, conn = box_mesh(*n, lX=100.0, lY=100.0, lZ=100.0)
WHAT = os.environ.get("EWB_WHAT", "le")  # le | vm | nh   (vm / nh: single-role kernel, EWB_NW=12)
if WHAT == "vm":
    asm = ElementAssembly("C3D8", conn, coords, "vonmises", [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0], box=n)
    import numpy as np
    G = 2.1e4 / (2 * 1.22)
    gmax = 2.0 * 355.0 / (np.sqrt(3.0) * G)
    d = 1e-6 * np.random.default_rng(0).standard_normal(asm.nDof)
    d[0::3] += 0.5 * gmax * coords[:, 1] ** 2 / 100.0
    dU = torch.as_tensor(d)
elif WHAT == "nh":
    asm = ElementAssembly("C3D8TL", conn, coords, "neohookewa", [91304.34783, 100000.0], box=n)
    dU = 1e-2 * torch.randn(asm.nDof, dtype=torch.float64)
else:
    asm = ElementAssembly("C3D8", conn, coords, "linearelastic", [2.1e4, 0.22], box=n)
    dU = 1e-3 * torch.randn(asm.nDof, dtype=torch.float64)
asm.U.copy_(dU)
asm.dU.copy_(dU)
for _ in range(3):
    asm.assemble()
asm.poll()
lib = asm.lib
lib.ewb_debug_timing.restype = C.c_int64
lib.ewb_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
NW = int(os.environ.get("EWB_NW", "12"))
buf = np.zeros(225 * 64 * NW * 12, dtype=np.int64)
m = lib.ewb_debug_timing(asm.plan, buf.ctypes.data, buf.size)
t = buf[:m].reshape(-1, NW, 12)
t = t[t[:, 0, 7] > 0]
names = ["phaseA | wait_producer", "wait_round", "elementBlocks", "emission", "wait_flush", "flush", "total", "steps", "A:loads+J,D", "A:inverse+rec", "A:material", "A:state store"]
print("CTAs", t.shape[0], "steps per CTA", np.unique(t[:, 0, 7]))
full = t[t[:, :, 6].max(axis=1) > np.percentile(t[:, :, 6].max(axis=1), 50)]  # the heavier half: full tiles
steps = full[:, :, 7].mean()
print("cycles per plane step, mean over warps of full-tile CTAs:")
for i, nme in [(j, names[j]) for j in (0, 8, 9, 10, 11, 1, 2, 3, 4, 5, 6)]:
    print(f"  {nme:14s} {full[:, :, i].mean() / steps:10.0f}   (per-warp min {full[:, :, i].mean(axis=0).min() / steps:8.0f}  max {full[:, :, i].mean(axis=0).max() / steps:8.0f})")

# per-warp table (warp w runs on scheduler w % 4): cycles per plane step
print("per-warp cycles per plane step (rows: warps; cols: wait_producer, wait_round, blocks, emission, wait_flush, flush, total)")
for w in range(NW):
    print(f"  warp {w:2d} smsp {w % 4}: " + " ".join(f"{full[:, w, i].mean() / steps:8.0f}" for i in (0, 1, 2, 3, 4, 5, 6)))
