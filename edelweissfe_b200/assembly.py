"""Device-resident element set: the B200 counterpart of the reference's
(elements dict + DofManager + VIJSystemMatrix + CSRGenerator) quartet for one element type.

PyTorch owns the buffers; all arithmetic happens in libedelweiss_b200.so (hand-written CUDA,
sm_100a) through the C ABI of include/edelweiss_b200.h.  No CPU fallback.

Reference interfaces mirrored (paths below /root/reference/edelweissfe/):
  numerics/dofmanager.py:445-471, 522-557     element dof lists, VIJ layout
  numerics/csrgenerator.pyx:47-115            CSR pattern + updateCSR
  solvers/nonlinearimplicitstatic.py:794-849  computeElements
  elements/displacementelement/element.py:373-379   acceptLastState
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import EwbBuffers, EwbError, check


class CutbackRequest(Exception):
    """Same meaning as edelweissfe.utils.exceptions.CutbackRequest(msg, cutbackSize)."""

    def __init__(self, message, cutbackSize):
        super().__init__(message)
        self.cutbackSize = cutbackSize


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def morton_order(coords: np.ndarray) -> np.ndarray:
    """Permutation of the nodes along a Z-order curve of their coordinates (10 bits per axis): spatially close nodes
    become neighbours in the visiting order of the row gather (ewb_plan_set_gather_order)."""
    x = np.asarray(coords, dtype=np.float64)
    lo, hi = x.min(axis=0), x.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((x - lo) / span * 1024.0).astype(np.uint64), 1023)

    def spread(v):  # 10 bits -> every third bit
        v = (v | (v << 16)) & np.uint64(0x030000FF)
        v = (v | (v << 8)) & np.uint64(0x0300F00F)
        v = (v | (v << 4)) & np.uint64(0x030C30C3)
        v = (v | (v << 2)) & np.uint64(0x09249249)
        return v

    key = (spread(q[:, 0]) << np.uint64(2)) | (spread(q[:, 1]) << np.uint64(1)) | spread(q[:, 2])
    return np.ascontiguousarray(np.argsort(key, kind="stable").astype(np.int32))


def pipeline_pieces(bounds, plane_dofs):
    """Dof ranges of the chunk pipeline (pure index logic; tests/test_pipeline_pieces.py).  bounds: node-plane boundaries of the
    fused kernel's x-chunks (ewb_plan_x_chunks).  Chunk c reads U / dU of the node planes [bounds[c] - 1, bounds[c + 1]] and writes
    P / F of [bounds[c], bounds[c + 1]).  Returns per chunk (up_lo, up_hi, down_lo, down_hi): the part of dU that has to be uploaded
    before chunk c can run and was not uploaded for an earlier chunk, and the part of P that is final once chunk c has run."""
    n_planes = bounds[-1]
    out = []
    for c in range(len(bounds) - 1):
        lo = 0 if c == 0 else (bounds[c] + 1) * plane_dofs
        hi = min(bounds[c + 1] + 1, n_planes) * plane_dofs
        out.append((lo, max(lo, hi), bounds[c] * plane_dofs, bounds[c + 1] * plane_dofs))
    return out


class ElementAssembly:
    def __init__(self, elType, conn, coords, material, props, device="cuda:0", box=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise EwbError("edelweissfe_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.elType = elType.upper()
        self.elCode = _lib.ELEMENT_CODES[self.elType]
        self.material = material.lower()
        self.matCode = _lib.MATERIAL_CODES[self.material]
        self.props = np.ascontiguousarray(np.asarray(props, dtype=np.float64))
        self.nn = _lib.ELEMENT_NODES[self.elCode]
        self.nGp = _lib.ELEMENT_GAUSS[self.elCode]
        self.nState = 12 + _lib.MATERIAL_NSTATE[self.matCode]
        conn = np.ascontiguousarray(np.asarray(conn, dtype=np.int32))
        assert conn.ndim == 2 and conn.shape[1] == self.nn
        coords_t = torch.as_tensor(np.asarray(coords, dtype=np.float64)) if not torch.is_tensor(coords) else coords
        self.nEl = conn.shape[0]
        self.nNode = coords_t.shape[0]
        self.nDof = 3 * self.nNode
        self.nDofEl = 3 * self.nn
        dev_index = self.device.index or 0
        plan = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ewb_plan_create(C.byref(plan), self.elCode, self.nEl, self.nNode, conn.ctypes.data_as(C.c_void_p), dev_index))
        self.plan = plan
        if box is not None:
            check(self.lib.ewb_plan_set_box(self.plan, int(box[0]), int(box[1]), int(box[2])))
        # arbitrary-mesh path: the row gather visits the nodes in Morton order of the coordinates (locality hint, results unchanged)
        # (measured on B200: DRAM reads of the gather 16.7 -> 9.7 GB for C3D20 100x100x50, but the kernel is latency bound and
        # does not get faster, so the hint is off unless EWB_GATHER_ORDER=morton)
        if os.environ.get("EWB_GATHER_ORDER", "none") == "morton":
            order = morton_order(coords_t.detach().cpu().numpy())
            check(self.lib.ewb_plan_set_gather_order(self.plan, order.ctypes.data_as(C.c_void_p)))
        # task-stream kernel for 20-node hexahedra (opt-in, EWB_STREAM=1): elements are processed along a Morton curve of their
        # centroids, so that a node's CSR rows are gathered soon after the matrices of its elements were written (locality hint,
        # results unchanged)
        if self.nn == 20 and os.environ.get("EWB_STREAM", "0") == "1" and os.environ.get("EWB_ELEMENT_ORDER", "morton") == "morton":
            cen = coords_t.detach().cpu().numpy()[conn.astype(np.int64)].mean(axis=1)
            check(self.lib.ewb_plan_set_element_order(self.plan, morton_order(cen).ctypes.data_as(C.c_void_p)))
        self.nnz = self.lib.ewb_plan_nnz(self.plan)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.coords = coords_t.to(**f64).contiguous()
        self.U = torch.zeros(self.nDof, **f64)
        self.dU = torch.zeros(self.nDof, **f64)
        self.P = torch.zeros(self.nDof, **f64)
        self.F = torch.zeros(self.nDof, **f64)
        self.state_ref = torch.zeros(self.nState, self.nEl, self.nGp, **f64)
        self.state_temp = torch.zeros(self.nState, self.nEl, self.nGp, **f64)
        self.csr_data = torch.zeros(self.nnz, **f64)
        self._pattern = None
        self._props_c = self.props.ctypes.data_as(C.POINTER(C.c_double))

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.ewb_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- pattern ---------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def csr_pattern(self):
        """(indptr, indices) int32 device tensors — CSRGenerator.__init__ (csrgenerator.pyx:47-98)."""
        if self._pattern is None:
            indptr = torch.empty(self.nDof + 1, dtype=torch.int32, device=self.device)
            indices = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
            check(self.lib.ewb_plan_csr_pattern(self.plan, _ptr(indptr), _ptr(indices), self._stream()))
            self._pattern = (indptr, indices)
        return self._pattern

    def slot_map(self, e0=0, e1=None):
        """COO->CSR slot map x for elements [e0, e1) (csrgenerator.pyx:82-98)."""
        e1 = self.nEl if e1 is None else e1
        x = torch.empty((e1 - e0) * self.nDofEl**2, dtype=torch.int32, device=self.device)
        check(self.lib.ewb_plan_slot_map(self.plan, e0, e1, _ptr(x), self._stream()))
        return x

    # ---- hot path ----------------------------------------------------------------------------
    def _buffers(self, vij=None):
        return EwbBuffers(_ptr(self.coords), _ptr(self.U), _ptr(self.dU), _ptr(self.state_ref), _ptr(self.state_temp),
                          _ptr(self.csr_data), _ptr(self.P), _ptr(self.F), _ptr(vij))

    def assemble(self, flags=0, vij=None, time=(0.0, 0.0), dT=0.0):
        """One computeElements + updateCSR pass on the device (asynchronous on the current stream).
        Results: self.csr_data, self.P, self.F, self.state_temp."""
        buf = self._buffers(vij)
        t = (C.c_double * 2)(*time)
        check(self.lib.ewb_assemble(self.plan, self.matCode, self._props_c, len(self.props), C.byref(buf), t, float(dT), int(flags), self._stream()))

    def poll(self):
        """Synchronise and turn a device-side material failure into CutbackRequest (vonmises.py:230-231)."""
        pnew = C.c_double(1.0)
        rc = check(self.lib.ewb_poll_status(self.plan, self._stream(), C.byref(pnew)))
        if rc == _lib.EWB_CUTBACK:
            raise CutbackRequest(self.lib.ewb_last_error().decode(), pnew.value)

    def compute_elements_vij(self, flags=0):
        """Reference-layout outputs: V (VIJ values), Pe [nEl, nDofEl]."""
        V = torch.empty(self.nEl * self.nDofEl**2, dtype=torch.float64, device=self.device)
        Pe = torch.empty(self.nEl, self.nDofEl, dtype=torch.float64, device=self.device)
        buf = self._buffers(V)
        check(self.lib.ewb_compute_elements_vij(self.plan, self.matCode, self._props_c, len(self.props), C.byref(buf), _ptr(Pe), int(flags), self._stream()))
        return V, Pe

    def update_csr(self, V):
        check(self.lib.ewb_update_csr(self.plan, _ptr(V), _ptr(self.csr_data), self._stream()))
        return self.csr_data

    def body_force(self, load, pext=None):
        """PExt += body-force load of every element (computeBodyForces, nonlinearimplicitstatic.py:516-557).
        pext: device tensor [nDof] (accumulated into); a zeroed one is created when omitted."""
        if pext is None:
            pext = torch.zeros(self.nDof, dtype=torch.float64, device=self.device)
        ld = (C.c_double * 3)(*[float(v) for v in load])
        check(self.lib.ewb_body_force(self.plan, _ptr(self.coords), ld, _ptr(pext), self._stream()))
        return pext

    def body_force_host(self, load):
        return self.body_force(load).cpu().numpy()

    def accept_last_state(self):
        """acceptLastState for every element (element.py:373-379): stateTemp becomes stateRef."""
        self.state_ref, self.state_temp = self.state_temp, self.state_ref

    def _dofs_dev(self, dofs):
        if torch.is_tensor(dofs) and dofs.dtype == torch.int32 and dofs.device == self.device:
            return dofs.contiguous()
        return torch.as_tensor(np.asarray(dofs, dtype=np.int32), device=self.device).contiguous()

    def apply_dirichlet_k(self, dofs):
        """NIST.applyDirichletK on the device (nonlinearimplicitstatic.py:559-593, pattern kept like mk2.pyx:66-109)."""
        d = self._dofs_dev(dofs)
        check(self.lib.ewb_apply_dirichlet_k(self.plan, _ptr(self.csr_data), _ptr(d), d.numel(), self._stream()))

    def apply_dirichlet_r(self, R, dofs, values=None):
        """NIST.applyDirichlet on a device residual (nonlinearimplicitstatic.py:595-623): R[dofs] = values (zeros when None)."""
        d = self._dofs_dev(dofs)
        v = None if values is None else torch.as_tensor(np.asarray(values, dtype=np.float64), device=self.device).contiguous()
        check(self.lib.ewb_apply_dirichlet_r(_ptr(R), _ptr(d), _ptr(v), d.numel(), self._stream()))
        return R

    def spmv(self, x, y=None):
        """y = K x with the assembled CSR values (device tensors)."""
        if y is None:
            y = torch.empty(self.nDof, dtype=torch.float64, device=self.device)
        check(self.lib.ewb_spmv(self.plan, _ptr(self.csr_data), _ptr(x), _ptr(y), self._stream()))
        return y

    def pcg_solve(self, b, dirichlet_dofs=None, rel_tol=1e-10, max_iter=20000, x=None):
        """NIST.linearSolve on the device (nonlinearimplicitstatic.py:727-751): Jacobi-PCG on the assembled matrix with the
        Dirichlet rows treated as identity rows (x[dirichlet] = b[dirichlet]).  b, x: device tensors [nDof].
        Returns (x, iterations, relative residual)."""
        if x is None:
            x = torch.empty(self.nDof, dtype=torch.float64, device=self.device)
        d = self._dofs_dev(dirichlet_dofs if dirichlet_dofs is not None else np.zeros(0, dtype=np.int32))
        it, rr = C.c_int(0), C.c_double(0.0)
        check(self.lib.ewb_pcg_solve(self.plan, _ptr(self.csr_data), _ptr(b), _ptr(x), _ptr(d) if d.numel() else C.c_void_p(0), d.numel(),
                                     float(rel_tol), int(max_iter), C.byref(it), C.byref(rr), self._stream()))
        return x, it.value, rr.value

    def pcg_solve_host(self, b, dirichlet_dofs=None, rel_tol=1e-10, max_iter=20000):
        """Host-facing form: b [nDof] in, x [nDof] out (numpy); the matrix stays on the device."""
        hb = self._pinned("b", self.nDof)
        hb.numpy()[:] = b
        bd = getattr(self, "_pcg_b", None)
        if bd is None:
            bd = self._pcg_b = torch.empty(self.nDof, dtype=torch.float64, device=self.device)
            self._pcg_x = torch.empty(self.nDof, dtype=torch.float64, device=self.device)
        bd.copy_(hb, non_blocking=True)
        x, it, rr = self.pcg_solve(bd, dirichlet_dofs, rel_tol, max_iter, x=self._pcg_x)
        hx = self._pinned("x", self.nDof)
        hx.copy_(x, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return hx.numpy().copy(), it, rr

    def surface_pressure(self, elems, faces, pressure, pext=None):
        """PExt += pressure load on the (element index, Abaqus face id) pairs (computeDistributedLoads, :460-508)."""
        if pext is None:
            pext = torch.zeros(self.nDof, dtype=torch.float64, device=self.device)
        e = np.ascontiguousarray(np.asarray(elems, dtype=np.int32))
        f = np.ascontiguousarray(np.asarray(faces, dtype=np.int32))
        check(self.lib.ewb_surface_pressure(self.plan, _ptr(self.coords), e.size, e.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                                            float(pressure), _ptr(pext), self._stream()))
        return pext

    def surface_pressure_host(self, elems, faces, pressure):
        return self.surface_pressure(elems, faces, pressure).cpu().numpy()

    # ---- state layout --------------------------------------------------------------------------
    def state_aos(self, which="temp"):
        """[nEl, nGp, nState] like the reference's per-element _stateVars* arrays."""
        src = self.state_temp if which == "temp" else self.state_ref
        out = torch.empty(self.nEl, self.nGp, self.nState, dtype=torch.float64, device=self.device)
        check(self.lib.ewb_state_to_aos(_ptr(src), _ptr(out), self.nEl, self.nGp, self.nState, self._stream()))
        return out

    def set_state_aos(self, aos, which="ref"):
        a = torch.as_tensor(aos, dtype=torch.float64).to(self.device).contiguous()
        assert tuple(a.shape) == (self.nEl, self.nGp, self.nState)
        dst = self.state_temp if which == "temp" else self.state_ref
        check(self.lib.ewb_state_to_soa(_ptr(a), _ptr(dst), self.nEl, self.nGp, self.nState, self._stream()))

    # ---- host-facing calls (what the solver plugin uses) -----------------------------------------
    def _pinned(self, name, n):
        buf = getattr(self, "_pin_" + name, None)
        if buf is None or buf.numel() != n:
            buf = torch.empty(n, dtype=torch.float64).pin_memory()
            setattr(self, "_pin_" + name, buf)
        return buf

    def compute_host(self, U, dU, stateRef_aos=None, stateTemp_aos=None, time=(0.0, 0.0), dT=0.0, flags=0):
        """computeElements with HOST arrays (what the solver plugin calls): U, dU [nDof] in, (P, F) out; K stays on the device
        (csr_data_host() / the device solver consume it).  Raises CutbackRequest.  The returned arrays are views of pinned buffers
        that the next call overwrites (the plugin adds them to the solver's vectors at once).

        Gauss-point state is DEVICE RESIDENT by default: the pass reads state_ref and writes state_temp on the device, and the
        caller commits with accept_last_state() / reads back with download_state_temp() when an increment is accepted — per Newton
        iteration only 2 + 2 dof vectors cross PCIe.  Passing stateRef_aos / stateTemp_aos ([nEl, nGp, nState], the reference's
        per-element layout) restores the stateless form: stateRef is uploaded before and stateTemp downloaded after the pass."""
        hU, hdU = self._pinned("U", self.nDof), self._pinned("dU", self.nDof)
        # U / dU None: the caller has filled the pinned input buffers of host_io() in place (no staging copy)
        if U is not None and dU is not None and self.nDof >= (1 << 20):
            # two large pageable vectors: stage them on two threads (NumPy copies release the GIL) and start the first upload early
            if getattr(self, "_stager", None) is None:
                from concurrent.futures import ThreadPoolExecutor

                self._stager = ThreadPoolExecutor(max_workers=1)
            fut = self._stager.submit(np.copyto, hdU.numpy(), np.asarray(dU))
            hU.numpy()[:] = U
            self.U.copy_(hU, non_blocking=True)
            fut.result()
            self.dU.copy_(hdU, non_blocking=True)
        else:
            if U is not None:
                hU.numpy()[:] = U
            if dU is not None:
                hdU.numpy()[:] = dU
            self.U.copy_(hU, non_blocking=True)
            self.dU.copy_(hdU, non_blocking=True)
        if stateRef_aos is not None:
            self.upload_state_ref(stateRef_aos)
        self.assemble(flags, time=time, dT=dT)
        hP, hF = self._pinned("P", self.nDof), self._pinned("F", self.nDof)
        hP.copy_(self.P, non_blocking=True)
        hF.copy_(self.F, non_blocking=True)
        if stateTemp_aos is not None:
            self.download_state_temp(stateTemp_aos, sync=False)
        self.poll()
        if stateTemp_aos is not None:
            np.asarray(stateTemp_aos).reshape(-1)[:] = self._pinned("S", self.nEl * self.nGp * self.nState).numpy()
        return hP.numpy(), hF.numpy()  # views of the pinned output buffers: valid until the next call

    # ---- the lean per-iteration call: U_n device resident, dU in, P and the flux norm out --------------------------------------
    def begin_increment(self, U_n=None):
        """U_n of the increment that starts (first argument of NIST.solveIncrement, nonlinearimplicitstatic.py:334-347) becomes
        device resident: every Newton iteration of the increment then uploads dU only and forms U_np = U_n + dU on the device —
        the same IEEE addition the solver does on the host (:416-417), bit for bit.  None: the pinned U buffer of host_io() was
        filled in place."""
        if getattr(self, "Un", None) is None:
            self.Un = torch.zeros(self.nDof, dtype=torch.float64, device=self.device)
            self._fsum = torch.zeros(1, dtype=torch.float64, device=self.device)
            self._pin_fsum = torch.zeros(1, dtype=torch.float64).pin_memory()
        hU = self._pinned("U", self.nDof)
        if U_n is not None:
            hU.numpy()[:] = U_n
        self.Un.copy_(hU, non_blocking=True)

    def _increment_upload(self, dU):
        if getattr(self, "Un", None) is None:
            raise EwbError("compute_host_increment needs begin_increment(U_n) first (U_n of the increment is device resident)")
        hdU = self._pinned("dU", self.nDof)
        if dU is not None:
            hdU.numpy()[:] = dU
        self.dU.copy_(hdU, non_blocking=True)
        torch.add(self.Un, self.dU, out=self.U)

    def _increment_download(self, n_owned=None):
        hP = self._pinned("P", self.nDof)
        hP.copy_(self.P, non_blocking=True)
        F = self.F if n_owned is None else self.F[:n_owned]
        torch.sum(F, dim=0, keepdim=True, out=self._fsum)  # F >= 0: the 1-norm checkConvergence needs (:785-790)
        self._pin_fsum.copy_(self._fsum, non_blocking=True)
        return hP

    def _registered_input(self, arr):
        """A pageable NumPy array the caller keeps re-using (the solver's dU lives for a whole step, nonlinearimplicitstatic.py:163)
        as a pinned torch tensor: the array's memory is registered with the driver once (cudaHostRegister) and unregistered when
        the array is garbage collected, so that its upload needs no staging copy.  None when that is not possible."""
        import weakref

        if os.environ.get("EWB_HOST_REGISTER", "1") != "1":
            return None
        arr = np.asarray(arr)
        if arr.dtype != np.float64 or not arr.flags["C_CONTIGUOUS"] or arr.nbytes < int(os.environ.get("EWB_HOST_REGISTER_MIN", 1 << 20)):
            return None
        reg = self.__dict__.setdefault("_registered", {})
        key = (arr.ctypes.data, arr.nbytes)
        if key not in reg:
            owner = arr
            while isinstance(getattr(owner, "base", None), np.ndarray):  # register / track the array that owns the memory
                owner = owner.base
            if owner.ctypes.data != arr.ctypes.data or owner.nbytes != arr.nbytes:
                return None
            lib = self.lib
            ptr = arr.ctypes.data
            if lib.ewb_host_register(C.c_void_p(ptr), arr.nbytes) != 0:
                return None

            def _release(ptr=ptr, key=key, reg=reg, lib=lib):
                reg.pop(key, None)
                try:
                    lib.ewb_host_unregister(C.c_void_p(ptr))
                except Exception:  # noqa: BLE001 - interpreter shutdown
                    pass

            try:
                weakref.finalize(owner, _release)
            except TypeError:
                lib.ewb_host_unregister(C.c_void_p(ptr))
                return None
            reg[key] = True
        t = torch.from_numpy(arr.reshape(-1))
        return t if t.is_pinned() else None

    def x_chunks(self, flags=0):
        """Node-plane boundaries of the fused kernel's independent x-chunks (ewb_plan_x_chunks), or None when this plan does not
        run the chunked kernel.  Cached per flags."""
        cache = self.__dict__.setdefault("_x_chunks", {})
        if flags not in cache:
            bounds = (C.c_int32 * 130)()
            n = check(self.lib.ewb_plan_x_chunks(self.plan, self.matCode, self._props_c, len(self.props), int(flags), bounds, 130))
            cache[flags] = None if n <= 0 else [int(bounds[i]) for i in range(n + 1)]
        return cache[flags]

    def _pipeline_streams(self):
        st = getattr(self, "_pipe_streams", None)
        if st is None:
            with torch.cuda.device(self.device):
                st = self._pipe_streams = dict(h2d=torch.cuda.Stream(), d2h=torch.cuda.Stream(), k=[torch.cuda.Stream(), torch.cuda.Stream()])
        return st

    def _increment_pipelined(self, dU, bounds, flags, launch=None, n_owned=None):
        """The chunk pipeline behind compute_host_increment: for every x-chunk c  [upload dU of its node planes, U = U_n + dU] ->
        [kernel of chunk c] -> [download P of its node planes], on a copy-in stream, two alternating kernel streams and a copy-out
        stream (PCIe is full duplex; kernels of different chunks are independent and fill the SMs together)."""
        hdU, hP = self._pinned("dU", self.nDof), self._pinned("P", self.nDof)
        hdU_np = hdU.numpy()
        dU = None if dU is None else np.asarray(dU)
        if dU is not None:
            reg = self._registered_input(dU)  # upload straight from the caller's array when it can be pinned in place
            if reg is not None:
                hdU, dU = reg, None
        st = self._pipeline_streams()
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for s in (st["h2d"], st["d2h"], *st["k"]):
            s.wait_event(start)
        nPlanes = bounds[-1]
        pd = self.nDof // nPlanes  # dofs per node plane
        buf = self._buffers()
        done = []
        for c, (lo, hi, a, b) in enumerate(pipeline_pieces(bounds, pd)):
            if dU is not None and hi > lo:
                hdU_np[lo:hi] = dU[lo:hi]  # pageable input: staged chunk by chunk, behind the previous chunk's upload and kernel
            with torch.cuda.stream(st["h2d"]):
                if hi > lo:
                    self.dU[lo:hi].copy_(hdU[lo:hi], non_blocking=True)
                    torch.add(self.Un[lo:hi], self.dU[lo:hi], out=self.U[lo:hi])
                up = torch.cuda.Event()
                up.record(st["h2d"])
            ks = st["k"][c & 1]
            ks.wait_event(up)
            if launch is None:
                check(self.lib.ewb_assemble_chunks(self.plan, self.matCode, self._props_c, len(self.props), C.byref(buf), int(flags), c, c + 1,
                                                   C.c_void_p(ks.cuda_stream)))
            else:
                launch(c, ks)
            kd = torch.cuda.Event()
            kd.record(ks)
            done.append(kd)
            if launch is None:
                st["d2h"].wait_event(kd)
                with torch.cuda.stream(st["d2h"]):
                    hP[a:b].copy_(self.P[a:b], non_blocking=True)
        for kd in done:
            cur.wait_event(kd)
        return hP

    def compute_host_increment_pipelined(self, dU, flags=0):
        """compute_host_increment with the transfers overlapped chunk by chunk (BoxGen plans on the fused kernel; falls back to the
        plain call otherwise).  Same results, bitwise."""
        bounds = self.x_chunks(flags)
        if bounds is None or len(bounds) < 3 or getattr(self, "Un", None) is None:
            return self.compute_host_increment(dU, flags=flags)
        hP = self._increment_pipelined(dU, bounds, flags)
        torch.sum(self.F, dim=0, keepdim=True, out=self._fsum)
        self._pin_fsum.copy_(self._fsum, non_blocking=True)
        torch.cuda.current_stream(self.device).wait_stream(self._pipe_streams["d2h"])
        self.poll()
        return hP.numpy(), float(self._pin_fsum[0])

    def compute_host_increment(self, dU, time=(0.0, 0.0), dT=0.0, flags=0):
        """One Newton iteration of the current increment with HOST dU (after begin_increment): returns (P, sum|F|).  P is a view of
        a pinned buffer (valid until the next call); F itself stays on the device (self.F) — the solver only ever takes its 1-norm
        per field (computeSpatialAveragedFluxes, :771-792).  Per iteration one dof vector crosses PCIe in each direction.
        Raises CutbackRequest."""
        self._increment_upload(dU)
        self.assemble(flags, time=time, dT=dT)
        hP = self._increment_download()
        self.poll()
        return hP.numpy(), float(self._pin_fsum[0])

    def host_io(self):
        """Pinned host buffers of compute_host as NumPy views: (U, dU) inputs a caller may fill in place, (P, F) outputs."""
        return tuple(self._pinned(k, self.nDof).numpy() for k in ("U", "dU", "P", "F"))

    def _aos_device_scratch(self):
        scratch = getattr(self, "_aos_scratch", None)
        if scratch is None:
            scratch = self._aos_scratch = torch.empty(self.nEl * self.nGp * self.nState, dtype=torch.float64, device=self.device)
        return scratch

    def upload_state_ref(self, stateRef_aos):
        """Host [nEl, nGp, nState] (the elements' _stateVarsRef, element.py:225-236) -> device state_ref (SoA)."""
        hS = self._pinned("S", self.nEl * self.nGp * self.nState)
        hS.numpy()[:] = np.asarray(stateRef_aos).reshape(-1)
        scratch = self._aos_device_scratch()
        scratch.copy_(hS, non_blocking=True)
        check(self.lib.ewb_state_to_soa(_ptr(scratch), _ptr(self.state_ref), self.nEl, self.nGp, self.nState, self._stream()))

    def download_state_temp(self, out=None, sync=True):
        """Device state_temp (SoA) -> host [nEl, nGp, nState]; `out` is filled in place when given."""
        hS = self._pinned("S", self.nEl * self.nGp * self.nState)
        scratch = self._aos_device_scratch()
        check(self.lib.ewb_state_to_aos(_ptr(self.state_temp), _ptr(scratch), self.nEl, self.nGp, self.nState, self._stream()))
        hS.copy_(scratch, non_blocking=True)
        if not sync:
            return None
        torch.cuda.current_stream(self.device).synchronize()
        res = hS.numpy().reshape(self.nEl, self.nGp, self.nState)
        if out is not None:
            np.asarray(out).reshape(self.nEl, self.nGp, self.nState)[...] = res
            return out
        return res.copy()

    def csr_pattern_host(self):
        indptr, indices = self.csr_pattern()
        return indptr.cpu().numpy(), indices.cpu().numpy()

    def csr_data_host(self):
        hK = self._pinned("K", self.nnz)
        hK.copy_(self.csr_data, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return hK.numpy()

    # ---- host round trips -----------------------------------------------------------------------
    def to_scipy(self):
        import scipy.sparse as sp

        indptr, indices = self.csr_pattern()
        return sp.csr_matrix((self.csr_data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(self.nDof, self.nDof))

    def launch_count(self):
        return int(self.lib.ewb_launch_count())
