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

Gauss-point state is device resident during the Newton iterations.  The elements' own `_stateVarsRef` /
`_stateVarsTemp` arrays (element.py:225-236) are never re-bound — field outputs cache raw pointers into them
(utils/elementresultcollector.pyx:83-97): when an increment is accepted (`model.advanceToTime`, models/femodel.py:245-258)
the plugin first downloads the converged state into every element's `_stateVarsTemp` IN PLACE and swaps the device
buffers, then the reference's own `acceptLastState` (element.py:373-379) copies it into `_stateVarsRef` as always.

With `*solver, solver=NISTB200, ..., b200solver=pcg` the matrix never leaves the device either: `assembleStiffnessCSR`
returns a device handle, `applyDirichletK` (:559-593) zeroes the rows on the device and `linearSolve` (:727-751) runs
the hand-written Jacobi-PCG; per iteration only dof-sized vectors cross PCIe.  The default (`b200solver=host`) hands a
scipy CSR matrix to whatever `linsolver=` names, like the reference.

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

    def gather_state_ref(self):
        """Contiguous [nEl, nInt, nState] copy of the elements' accepted state (element.py:225-236)."""
        return np.array([np.asarray(e._stateVarsRef) for e in self.elements])

    def scatter_state_temp(self, stateTemp):
        """Write the converged state into every element's OWN _stateVarsTemp array, in place (pointers held by field
        outputs stay valid); the reference's acceptLastState then copies it into _stateVarsRef (element.py:373-379)."""
        for e, st in zip(self.elements, stateTemp):
            e._stateVarsTemp[...] = st

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

    class DeviceCSR:
        """What assembleStiffnessCSR returns with b200solver=pcg: the assembled matrix as a device handle (no values on the host)."""

        def __init__(self, asm):
            self.asm = asm
            self.dirichlet = None
            self.shape = (asm.nDof, asm.nDof)

        def to_scipy(self):
            import scipy.sparse as sp

            indptr, indices = self.asm.csr_pattern_host()
            return sp.csr_matrix((self.asm.csr_data_host().copy(), indices, indptr), shape=self.shape)

    class NISTB200(NIST):
        identification = "NISTB200Solver"
        # b200solver: "host" = scipy CSR for the reference's linsolver (default), "pcg" = matrix stays on the device (Jacobi-PCG)
        # b200io: "lean" = U_n device resident per increment, per Newton iteration dU in and P + the flux 1-norm out (one dof vector each
        #         way); "full" = U_np, dU in and P, F out, the literal signature of NIST.computeElements
        NISTOptions = dict(NIST.NISTOptions, b200solver="host", b200pcgtol=1e-12, b200pcgmaxiter=100000, b200io="lean")

        def _b200_setup(self, elements):
            self._b200_dm = self.theDofManager
            ex = ElementSetExtraction(elements, self.theDofManager)
            self._b200_ex = ex
            self._b200_asm = backend_factory(ex.elType, ex.conn, ex.coords, ex.material, ex.props, box=ex.detect_box())
            self._b200_asm.upload_state_ref(ex.gather_state_ref())  # the device copy is authoritative until the step ends
            self._b200_map = None
            self._b200_index = {id(e): k for k, e in enumerate(ex.elements)}
            self.journal.message(
                f"B200 element loop: {len(ex.elements)} x {ex.elType} / {ex.material}, "
                f"{'fused BoxGen sweep' if ex.detect_box() else 'generic two-phase'} path", self.identification, 0)

        def solveStep(self, step, model, fieldOutputController, outputmanagers):
            """NIST.solveStep (:110-332) with the state commit hooked into model.advanceToTime (:296, models/femodel.py:245-258)."""
            original = model.advanceToTime

            def advanceToTime(time):
                asm = getattr(self, "_b200_asm", None)
                if asm is not None and getattr(self, "_b200_dm", None) is self.theDofManager:
                    self._b200_ex.scatter_state_temp(asm.download_state_temp())  # in place: field-output pointers stay valid
                    asm.accept_last_state()
                return original(time)

            model.advanceToTime = advanceToTime
            try:
                return super().solveStep(step, model, fieldOutputController, outputmanagers)
            finally:
                del model.advanceToTime

        def solveIncrement(self, U_n, dU, P, K, stepActions, model, timeStep, prevTimeStep, extrapolation, maxIter, maxGrowingIter):
            """NIST.solveIncrement (:334-456); remembers U_n so that the element loop can keep it on the device for the whole increment."""
            self._b200_Un = U_n
            self._b200_Un_sent = False
            self._b200_flux = None
            try:
                return super().solveIncrement(U_n, dU, P, K, stepActions, model, timeStep, prevTimeStep, extrapolation, maxIter, maxGrowingIter)
            finally:
                self._b200_Un = None

        def computeElements(self, elements, U_np, dU, P, K, F, timeStep):
            tic = _time.time()
            if getattr(self, "_b200_dm", None) is not self.theDofManager:
                self._b200_setup(elements)
            asm = self._b200_asm
            lean = str(self.options.get("b200io", "lean")).lower() == "lean" and getattr(self, "_b200_Un", None) is not None \
                and hasattr(asm, "compute_host_increment")
            tm = dict(time=(timeStep.stepTime, timeStep.totalTime), dT=timeStep.timeIncrement)
            try:
                if lean:
                    if not self._b200_Un_sent:
                        asm.begin_increment(np.asarray(self._b200_Un))
                        self._b200_Un_sent = True
                    # U_np = U_n + dU on the device (:416-417); BoxGen plans overlap the transfers with the kernel chunk by chunk
                    if hasattr(asm, "compute_host_increment_pipelined"):
                        Pel, self._b200_flux = asm.compute_host_increment_pipelined(np.asarray(dU))
                    else:
                        Pel, self._b200_flux = asm.compute_host_increment(np.asarray(dU), **tm)
                else:
                    Pel, Fel = asm.compute_host(np.asarray(U_np), np.asarray(dU), **tm)
                    self._b200_flux = None
            except Exception as e:  # device-side material failure -> the reference's cut-back request
                if type(e).__name__ == "CutbackRequest":
                    raise CutbackRequest(str(e), e.cutbackSize)
                raise
            P += Pel  # P[el] += Pe for all elements (:843)
            if not lean:
                F += Fel  # F[el] += abs(Pe)            (:844)
            self.computationTimes["elements"] += _time.time() - tic
            return P, K, F

        def computeSpatialAveragedFluxes(self, F):
            """:771-792.  In lean mode the element fluxes were summed on the device (F >= 0, so the sum is the 1-norm); whatever else
            the host vector F holds is added to it."""
            flux = getattr(self, "_b200_flux", None)
            if flux is None:
                return super().computeSpatialAveragedFluxes(F)
            out = dict.fromkeys(self.theDofManager.idcsOfFieldsInDofVector, 0.0)
            for field, nDof in self.theDofManager.nAccumulatedNodalFluxesFieldwise.items():
                host = float(np.linalg.norm(F[self.theDofManager.idcsOfFieldsInDofVector[field]], 1))
                out[field] = max(1e-10, (flux + host) / nDof)
            return out

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

        def computeDistributedLoads(self, distributedLoads, U_np, PExt, K, timeStep):
            """`type=pressure` on faces of 8-node hexahedra runs on the device (:460-508; the reference's Python element raises for
            every distributed load, element.py:255-288, so this is what makes BASELINE config 1 runnable with provider=edelweiss);
            anything else goes to the reference loop."""
            tic = _time.time()
            rest = []
            for dLoad in distributedLoads:
                asm = getattr(self, "_b200_asm", None)
                idx = getattr(self, "_b200_index", {})
                ok = asm is not None and hasattr(asm, "surface_pressure_host") and dLoad.loadType == "pressure" and self._b200_ex.conn.shape[1] == 8 \
                    and all(id(el) in idx for els in dLoad.surface.values() for el in els)
                if not ok:
                    rest.append(dLoad)
                    continue
                elems = [idx[id(el)] for els in dLoad.surface.values() for el in els]
                faces = [int(faceID) for faceID, els in dLoad.surface.items() for _ in els]
                load = np.atleast_1d(np.asarray(dLoad.getCurrentLoad(timeStep), dtype=float))
                PExt += asm.surface_pressure_host(elems, faces, float(load[0]))
            self.computationTimes["distributed loads"] += _time.time() - tic
            if rest:
                return super().computeDistributedLoads(rest, U_np, PExt, K, timeStep)
            return PExt, K

        def _b200_device_solver(self):
            return str(self.options.get("b200solver", "host")).lower() == "pcg" and hasattr(self._b200_asm, "pcg_solve_host")

        def assembleStiffnessCSR(self, K):
            tic = _time.time()
            asm = self._b200_asm
            if self._b200_device_solver():
                if np.any(np.asarray(K)):
                    raise NotImplementedError("b200solver=pcg: constraints that write into the system matrix need b200solver=host")
                self.computationTimes["CSR generation"] += _time.time() - tic
                return DeviceCSR(asm)
            KCsr = self.csrGenerator.updateCSR(K)  # whatever constraints wrote into the VIJ (elements left it zero)
            if self._b200_map is None:
                indptr, indices = asm.csr_pattern_host()
                self._b200_map = _pattern_map(indptr, indices, KCsr.indptr, KCsr.indices)
            KCsr.data[self._b200_map] += asm.csr_data_host()
            self.computationTimes["CSR generation"] += _time.time() - tic
            return KCsr

        def applyDirichletK(self, K, dirichlets):
            if not isinstance(K, DeviceCSR):
                return super().applyDirichletK(K, dirichlets)
            tic = _time.time()
            if dirichlets:
                K.dirichlet = np.concatenate([np.asarray(self.findDirichletIndices(d)).ravel() for d in dirichlets]).astype(np.int32)
                K.asm.apply_dirichlet_k(K.dirichlet)  # rows zeroed, 1 on the diagonal, on the device (:559-593)
            self.computationTimes["dirichlet K"] += _time.time() - tic
            return K

        def linearSolve(self, A, b):
            if not isinstance(A, DeviceCSR):
                return super().linearSolve(A, b)
            from edelweissfe.utils.exceptions import DivergingSolution

            tic = _time.time()
            ddU, iters, relres = A.asm.pcg_solve_host(np.asarray(b), A.dirichlet, float(self.options["b200pcgtol"]), int(self.options["b200pcgmaxiter"]))
            self.computationTimes["linear solve"] += _time.time() - tic
            if np.isnan(ddU).any() or relres > 1e3 * float(self.options["b200pcgtol"]):
                raise DivergingSolution("device PCG did not converge (%d iterations, relative residual %.2e)" % (iters, relres))
            return ddU

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
