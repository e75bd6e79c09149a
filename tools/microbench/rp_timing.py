"""Per-role cycle counters of the row-pipelined kernel (debug build: make -C edelweissfe_b200/csrc timing).
usage: EWB_KERNEL=rp484 python tools/microbench/rp_timing.py [le|vm|nh]"""
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

what = sys.argv[1] if len(sys.argv) > 1 else "le"
variant = os.environ.get("EWB_KERNEL", "rp4_4_4")
npw, ntw, ngw = (int("".join(ch for ch in c if ch.isdigit())) for c in variant.split("_")[-3:])
n = (100, 100, 100)
coords, conn = box_mesh(*n, lX=100.0, lY=100.0, lZ=100.0)
if what == "vm":
    asm = ElementAssembly("C3D8", conn, coords, "vonmises", [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0], box=n)
    G = 2.1e4 / (2 * 1.22)
    gmax = 2.0 * 355.0 / (np.sqrt(3.0) * G)
    d = 1e-6 * np.random.default_rng(0).standard_normal(asm.nDof)
    d[0::3] += 0.5 * gmax * coords[:, 1] ** 2 / 100.0
    dU = torch.as_tensor(d)
elif what == "nh":
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
NW = npw + ntw + ngw
buf = np.zeros(4096 * NW * 8, dtype=np.int64)
m = lib.ewb_debug_timing(asm.plan, buf.ctypes.data, buf.size)
t = buf[:m].reshape(-1, NW, 8)
t = t[t[:, 0, 6] > 0]
tot = t[:, :, 6].max(axis=1)
full = t[tot > np.percentile(tot, 50)]
print(f"{variant} {what}: CTAs {t.shape[0]}, CTA duration cycles median {np.median(tot):.0f} max {tot.max():.0f}")
roles = [("P", 0, npw, ["wait recEmpty", "wait cp.async", "issue", "phaseA+arrive"]), ("T", npw, npw + ntw, ["wait recFull", "wait slotEmpty", "tiles", "store+arrive"]),
         ("G", npw + ntw, NW, ["wait slotFull", "lean loads+adds", "lean stores", "other/general", "residual pass", "arrive"])]
for name, a, b, labels in roles:
    w = full[:, a:b, :]
    cnt = w[:, :, 7].mean()
    print(f" role {name}: items per warp {cnt:.0f}; cycles per item: " + ", ".join(f"{lab} {w[:, :, i].mean() / cnt:.0f}" for i, lab in zip(range(6), labels) if lab != "-")
          + f"; total/item {w[:, :, 6].mean() / cnt:.0f}")

if os.environ.get("EWB_TIMING_DETAIL"):
    order = np.argsort(-tot)
    print("slowest / fastest CTAs (index in launch order, total cycles; G-role: wait, lean loads, lean stores, general, residual):")
    idx_all = np.nonzero(buf[:m].reshape(-1, NW, 8)[:, 0, 6] > 0)[0]
    for k in list(order[:6]) + list(order[-4:]):
        g = t[k, npw + ntw:, :].mean(axis=0)
        print(f"  cta {idx_all[k]:4d} total {tot[k]:9.0f}  rows {t[k, -1, 7]:5.0f}  G per row: wait {g[0] / g[7]:6.0f} loads {g[1] / g[7]:6.0f} stores {g[2] / g[7]:6.0f} general {g[3] / g[7]:6.0f} resid {g[4] / g[7]:6.0f}")
