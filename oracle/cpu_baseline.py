"""CPU baseline legs of bench.py (TEST/BENCH INFRASTRUCTURE, not product), on a bounded BoxGen sample of the bench workload:

  run_reference()  kind "reference": the UNMODIFIED reference — its own element objects, the serial loop of
                   NIST.computeElements (solvers/nonlinearimplicitstatic.py:837-844) and CSRGenerator.updateCSR
                   (numerics/csrgenerator.pyx:100-115) — imported from baseline/_ref (tools/install_reference.py), 1 core
                   (the Python elements are GIL-bound and NISTParallel is wrong for them).
  run()            kind "port": the C/OpenMP restatement (oracle/element_loop.c) on every host core, or the NumPy port; a
                   second, clearly labelled figure ("not the reference") and the fallback when baseline/_ref is absent."""
import os
import time

import numpy as np

from . import port

_WL = {
    "boxgen100_c3d8_linearelastic": ("C3D8", "linearelastic", [2.1e4, 0.22], 1e-3),
    "boxgen200x100x100_c3d8_vonmises": ("C3D8", "vonmises", [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0], 5e-3),
    "boxgen100_c3d8tl_neohookewa": ("C3D8TL", "neohookewa", [91304.34783, 100000.0], 1e-2),
    "boxgen100x100x50_c3d20_linearelastic": ("C3D20", "linearelastic", [2.1e4, 0.22], 1e-3),
    "boxgen200_c3d8tl_neohookewa": ("C3D8TL", "neohookewa", [91304.34783, 100000.0], 1e-2),
}
# measured single-core cost of the reference per element, used only to size the sample for a time budget
_REF_MS_PER_ELEMENT = {"C3D8": 0.45, "C3D8TL": 2.0, "C3D20": 1.25}


def reference_available():
    try:
        import os as _os
        import sys as _sys

        root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
        if root not in _sys.path:
            _sys.path.insert(0, root)
        from tools import refshim

        return _os.path.isdir(_os.path.join(refshim._INSTALLED, "edelweissfe"))
    except Exception:
        return False


def run_reference(workload, steps=1, warmup=0, budget_s=60.0, sample_n=None):
    """Times the reference's own serial element loop + updateCSR (1 core).  The sample box is sized so that
    (steps + warmup) passes take about budget_s; capped at 20^3 (C3D8 / C3D8TL) or 10^3 (C3D20), BASELINE.md §3."""
    from tools import refdriver

    elType, material, props, scale = _WL[workload]
    cap = 10 if "20" in elType else 20
    if sample_n is None:
        per = budget_s / max(1, steps + warmup) / (_REF_MS_PER_ELEMENT[elType] * 1e-3)
        sample_n = int(max(3, min(cap, round(per ** (1.0 / 3.0)))))
    n = sample_n
    t0 = time.perf_counter()
    ref = refdriver.RefModel(elType, material, props, box=dict(nX=n, nY=n, nZ=n, lX=float(n), lY=float(n), lZ=float(n)))
    setup = time.perf_counter() - t0
    nDof = ref.dm.nDof
    rng = np.random.default_rng(0)
    dU = scale * rng.standard_normal(nDof)
    for _ in range(warmup):
        ref.assemble(dU, dU)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.assemble(dU, dU)
    dt = (time.perf_counter() - t0) / steps
    nEl = len(ref.elements)
    return dict(value=nEl / dt / 1e6, ms_per_step=dt * 1e3, cores=1, kind="reference", steps=steps, warmup=warmup,
                sample=f"BoxGen {n}x{n}x{n} {elType} {material} ({nEl} elements): the unmodified reference's element objects, serial "
                       f"NIST.computeElements loop + CSRGenerator.updateCSR, 1 core (model set-up {setup:.1f} s not timed); "
                       f"host has {os.cpu_count()} logical cores")


def run(workload, sample_n, steps=1, warmup=0):
    elType, material, props, scale = _WL[workload]
    nn = 20 if "20" in elType else 8
    if nn == 20:
        sample_n = max(4, sample_n // 3)
    n = sample_n
    coords, conn = port.boxgen(n, n, n, float(n), float(n), float(n), nnodes=nn)
    rng = np.random.default_rng(0)
    dU = scale * rng.standard_normal(3 * coords.shape[0])
    nGp = 27 if nn == 20 else 8
    state = np.zeros((conn.shape[0], nGp, 12 + port.MATERIAL_NSTATE[material]))
    try:
        from . import cport

        impl = cport.load()
    except Exception:
        impl = None
    if impl is not None and impl.supports(elType, material):
        # every host thread this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        runner = impl.make_runner(elType, material, props, coords, conn, dU, dU, state, nthreads=cores)
        kind_note = "C/OpenMP restatement of the reference algorithm (oracle/element_loop.c)"
    else:
        # pattern once (reference: per step), then the loop body per timed step
        dofs = port.element_dofs(conn)
        I, J = port.vij_pattern(dofs)  # noqa: E741
        indptr, indices, x = port.csr_pattern(I, J, 3 * coords.shape[0])

        def runner():
            Ke, Pe, st, failed = port.compute_elements(elType, material, props, coords, conn, dU, dU, state)
            port.update_csr(x, Ke.reshape(-1), indices.size)
            np.bincount(dofs.reshape(-1), weights=Pe.reshape(-1), minlength=3 * coords.shape[0])

        cores, kind_note = 1, "NumPy restatement (oracle/port.py), batched einsum"
    for _ in range(warmup):
        runner()
    t0 = time.perf_counter()
    for _ in range(steps):
        runner()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=conn.shape[0] / dt / 1e6, ms_per_step=dt * 1e3, cores=cores, kind="port", steps=steps, warmup=warmup,
                sample=f"BoxGen {n}x{n}x{n} {elType} {material} ({conn.shape[0]} elements), one computeElements+updateCSR pass; {kind_note}; "
                       f"host has {os.cpu_count()} logical cores")
