"""NISTB200 — the drop-in solver plugin: EdelweissFE's own NIST Newton-Raphson solver with the element
loop and the CSR assembly running on a B200.

Selection from an input file, after `import edelweissfe_b200.nistb200 as n; n.register()`:

    *solver, solver=NISTB200, name=theSolver
    *step, solver=theSolver, ...

`register()` adds `solverLibrary["NISTB200"] = "nistb200"` (config/solvers.py:48-53) and makes
`edelweissfe.solvers.nistb200` importable, which is how `getSolverByName` (config/solvers.py:56-76)
resolves solver names.  Everything else of the reference is untouched: input file, model tree, sections,
materials, step actions, constraints, linear solvers and outputs keep their interfaces.

Overridden members of NIST (solvers/nonlinearimplicitstatic.py):
    computeElements(elements, U_np, dU, P, K, F, timeStep) -> (P, K, F)       (:794-849)
    assembleStiffnessCSR(K) -> scipy.sparse.csr_matrix                        (:753-769)
The element contributions never pass through the VIJ triple: K (VIJ) keeps only what constraints write
into it afterwards (:889-896); assembleStiffnessCSR merges both into the reference's CSR pattern.

Gauss-point state stays visible to the reference: every element's `_stateVarsRef` / `_stateVarsTemp`
(element.py:225-236) are re-bound to views of two contiguous host arrays, so `acceptLastState`
(element.py:373-379), field outputs and `getResultArray` work unchanged.

This module needs the reference package to be importable; it is not used by bench.py or the GPU tests
(the GPU box has no reference tree).  Host-side logic is tested on CPU in tests/test_plugin_reference.py.
"""
from __future__ import annotations

import sys
import time as _time
import types

import numpy as np

_MATERIAL_BY_CLASS = {
    "LinearElasticMaterial": ("linearelastic", lambda m: [m._E, m._v]),
    "VonMisesMaterial": ("vonmises", lambda m: [m._E, m._v, m.yieldStress, m.HLin, m.deltaYieldStress, m.delta]),
    "NeoHookeanWaMaterial": ("neohookewa", lambda m: [m._mu, m._K]),
    "NeoHookeanWbMaterial": ("neohookewb", lambda m: [m._mu, m._K]),
    "NeoHookeanWcMaterial": ("neohookewc", lambda m: [m._mu, m._K]),
}
_SUPPORTED_ELTYPES = {"C3D8", "C3D8N", "C3D20", "C3D20N", "C3D8TL", "C3D8NTL", "C3D8R", "C3D8E", "C3D20R"}


def default_backend(elType, conn, coords, material, props, box=None):
    """The CUDA backend (no CPU fallback: raises when the library or a GPU is missing)."""
    from .assembly import ElementAssembly

    return ElementAssembly(elType, conn, coords, material, props, box=box)


class ElementSetExtraction:
    """Array view of `elements` (dict label -> reference element object) for the device path."""

    def __init__(self, elements, dofManager):
        els = list(elements.values())
        if not els:
            raise NotImplementedError("NISTB200 needs at least one element")
        types_ = {type(e).__name__ for e in els}
        if not types_ <= {"DisplacementElement", "DisplacementTLElement"}:
            raise NotImplementedError(f"NISTB200 supports provider 'edelweiss' displacement elements only, got {types_}")
        self.elements = els
        nn = els[0].nNodes
        nInt = els[0]._nInt
        if any(e.nNodes != nn or e._nInt != nInt for e in els):
            raise NotImplementedError("NISTB200 needs one element type per model")
        mats = {id(e.material) for e in els}
        if len(mats) != 1:
            raise NotImplementedError("NISTB200 needs one section/material for all elements")
        mat = els[0].material
        if type(mat).__name__ not in _MATERIAL_BY_CLASS:
            raise NotImplementedError(f"material {type(mat).__name__} is not implemented on the device")
        self.material, getprops = _MATERIAL_BY_CLASS[type(mat).__name__]
        self.props = [float(p) for p in getprops(mat)]
        tl = type(els[0]).__name__ == "DisplacementTLElement"
        self.elType = {(8, 8, False): "C3D8", (20, 27, False): "C3D20", (8, 8, True): "C3D8TL", (8, 1, False): "C3D8R", (8, 27, False): "C3D8E",
                       (20, 8, False): "C3D20R"}.get((nn, nInt, tl))
        if self.elType is None:
            raise NotImplementedError(f"element with {nn} nodes / {nInt} Gauss points (TL={tl}) is not implemented on the device")
        # element dof lists straight from the reference's DofManager (numerics/dofmanager.py:445-471)
        dofs = np.array([dofManager.idcsOfElementsInDofVector[e] for e in els], dtype=np.int64)
        if dofs.shape[1] != 3 * nn or not (np.array_equal(dofs[:, 1::3], dofs[:, 0::3] + 1) and np.array_equal(dofs[:, 2::3], dofs[:, 0::3] + 2)):
            raise NotImplementedError("NISTB200 needs three consecutive displacement dofs per node")
        if (dofs[:, 0::3] % 3).any() or dofManager.nDof % 3:
            raise NotImplementedError("NISTB200 needs a displacement-only dof vector (dof = 3*node + c)")
        self.conn = (dofs[:, 0::3] // 3).astype(np.int32)
        self.nNode = dofManager.nDof // 3
        coords = np.zeros((self.nNode, 3))
        for e, c in zip(els, self.conn):
            for n, i in zip(e.nodes, c):
                coords[i] = n.coordinates
        self.coords = coords
        self.nState = els[0]._stateVarsRef.shape[1]
        self.nInt = nInt
        # one contiguous host image of the Gauss-point state; the elements keep working on views of it
        self.stateRef = np.array([np.asarray(e._stateVarsRef) for e in els])
        self.stateTemp = np.array([np.asarray(e._stateVarsTemp) for e in els]) if np.shape(els[0]._stateVarsTemp) == np.shape(els[0]._stateVarsRef) else self.stateRef.copy()
        for k, e in enumerate(els):
            e._stateVarsRef = self.stateRef[k]
            e._stateVarsTemp = self.stateTemp[k]
            for i in range(nInt):
                sv = e._stateVars[i]
                sv["stress"] = self.stateRef[k, i, 0:6]
                sv["strain"] = self.stateRef[k, i, 6:12]
                sv["materialstate"] = self.stateRef[k, i, 12:]

    def detect_box(self):
        """(nX, nY, nZ) if the connectivity is BoxGen-ordered (generators/boxgen.py:168-185), else None."""
        if self.conn.shape[1] != 8:
            return None
        c0 = self.conn[0]
        nzs = int(c0[4] - c0[0])  # NZ = nZ + 1
        nyz = int(c0[3] - c0[0])  # NY * NZ
        if nzs < 2 or nyz < 2 * nzs or nyz % nzs:
            return None
        NZ, NY = nzs, nyz // nzs
        if self.nNode % (NY * NZ):
            return None
        NX = self.nNode // (NY * NZ)
        n = (NX - 1, NY - 1, NZ - 1)
        if min(n) < 1 or n[0] * n[1] * n[2] != self.conn.shape[0]:
            return None
        from .boxgen import box_mesh

        _, conn = box_mesh(*n)
        return n if np.array_equal(conn, self.conn) else None


def make_solver_class(backend_factory=default_backend):
    """Build NISTB200 on top of the reference's NIST (import deferred: needs the reference package)."""
    from edelweissfe.solvers.nonlinearimplicitstatic import NIST
    from edelweissfe.utils.exceptions import CutbackRequest

    class NISTB200(NIST):
        identification = "NISTB200Solver"

        def _b200_setup(self, elements):
            self._b200_dm = self.theDofManager
            ex = ElementSetExtraction(elements, self.theDofManager)
            self._b200_ex = ex
            self._b200_asm = backend_factory(ex.elType, ex.conn, ex.coords, ex.material, ex.props, box=ex.detect_box())
            self._b200_map = None
            self.journal.message(
                f"B200 element loop: {len(ex.elements)} x {ex.elType} / {ex.material}, "
                f"{'fused BoxGen sweep' if ex.detect_box() else 'generic two-phase'} path", self.identification, 0)

        def computeElements(self, elements, U_np, dU, P, K, F, timeStep):
            tic = _time.time()
            if getattr(self, "_b200_dm", None) is not self.theDofManager:
                self._b200_setup(elements)
            ex, asm = self._b200_ex, self._b200_asm
            try:
                Pel, Fel = asm.compute_host(np.asarray(U_np), np.asarray(dU), ex.stateRef, ex.stateTemp,
                                            time=(timeStep.stepTime, timeStep.totalTime), dT=timeStep.timeIncrement)
            except Exception as e:  # device-side material failure -> the reference's cut-back request
                if type(e).__name__ == "CutbackRequest":
                    raise CutbackRequest(str(e), e.cutbackSize)
                raise
            P += Pel  # P[el] += Pe for all elements (:843)
            F += Fel  # F[el] += abs(Pe)            (:844)
            self.computationTimes["elements"] += _time.time() - tic
            return P, K, F

        def computeBodyForces(self, bodyForces, U_np, PExt, K, timeStep):
            """Body forces acting on ALL elements run on the device (:516-557); partial element sets stay on the host loop."""
            tic = _time.time()
            rest = []
            for bForce in bodyForces:
                asm = getattr(self, "_b200_asm", None)
                if asm is not None and hasattr(asm, "body_force_host") and len(bForce.elementSet) == len(self._b200_ex.elements) \
                        and all(a is b for a, b in zip(bForce.elementSet, self._b200_ex.elements)):
                    PExt += asm.body_force_host(np.asarray(bForce.getCurrentLoad(timeStep), dtype=float))
                else:
                    rest.append(bForce)
            self.computationTimes["body forces"] += _time.time() - tic
            if rest:
                return super().computeBodyForces(rest, U_np, PExt, K, timeStep)
            return PExt, K

        def assembleStiffnessCSR(self, K):
            tic = _time.time()
            KCsr = self.csrGenerator.updateCSR(K)  # whatever constraints wrote into the VIJ (elements left it zero)
            asm = self._b200_asm
            if self._b200_map is None:
                indptr, indices = asm.csr_pattern_host()
                self._b200_map = _pattern_map(indptr, indices, KCsr.indptr, KCsr.indices)
            KCsr.data[self._b200_map] += asm.csr_data_host()
            self.computationTimes["CSR generation"] += _time.time() - tic
            return KCsr

    return NISTB200


def _pattern_map(indptr, indices, refIndptr, refIndices):
    """Position of every entry of the element CSR pattern inside the reference CSRGenerator's pattern
    (both canonical; identical when the model has no constraints)."""
    if indptr.shape == refIndptr.shape and indices.shape == refIndices.shape and np.array_equal(indptr, refIndptr) and np.array_equal(indices, refIndices):
        return np.arange(indices.size)
    out = np.empty(indices.size, dtype=np.int64)
    for r in range(indptr.size - 1):
        a0, a1 = indptr[r], indptr[r + 1]
        b0, b1 = refIndptr[r], refIndptr[r + 1]
        pos = np.searchsorted(refIndices[b0:b1], indices[a0:a1])
        if (pos >= b1 - b0).any() or not np.array_equal(refIndices[b0:b1][pos], indices[a0:a1]):
            raise RuntimeError("element CSR pattern is not contained in the solver's pattern")
        out[a0:a1] = b0 + pos
    return out


def register(backend_factory=default_backend):
    """Make `*solver, solver=NISTB200` resolvable by the reference (config/solvers.py:48-76)."""
    import edelweissfe.config.solvers as cfg

    mod = types.ModuleType("edelweissfe.solvers.nistb200")
    mod.NISTB200 = make_solver_class(backend_factory)
    sys.modules["edelweissfe.solvers.nistb200"] = mod
    import edelweissfe.solvers as pkg

    pkg.nistb200 = mod
    cfg.solverLibrary["NISTB200"] = "nistb200"
    return mod.NISTB200
