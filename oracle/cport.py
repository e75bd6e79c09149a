"""ctypes wrapper of oracle/element_loop.c (TEST / BENCH INFRASTRUCTURE ONLY)."""
import ctypes as C
import os

import numpy as np

from . import port

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libewb_oracle.so")
_EL = {"C3D8": 0, "C3D20": 1, "C3D8TL": 2, "C3D8R": 3, "C3D8E": 4, "C3D20R": 5}
_MAT = {"linearelastic": 0, "vonmises": 1, "neohookewa": 2, "neohookewb": 3, "neohookewc": 4}
_inst = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    def __init__(self):
        if not os.path.exists(_SO):
            raise FileNotFoundError(_SO + " (run `make -C oracle`)")
        lib = C.CDLL(_SO)
        lib.ewo_threads.restype = C.c_int
        lib.ewo_compute_elements.restype = C.c_int
        lib.ewo_compute_elements.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 8 + [C.c_int]
        lib.ewo_scatter_pf.restype = None
        lib.ewo_scatter_pf.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 4
        lib.ewo_update_csr.restype = None
        lib.ewo_update_csr.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        self.lib = lib

    def threads(self):
        return int(self.lib.ewo_threads())

    @staticmethod
    def supports(elType, material):
        tl = elType.upper() == "C3D8TL"
        return elType.upper() in _EL and material.lower() in _MAT and (tl or not material.lower().startswith("neohooke"))

    def compute_elements(self, elType, material, props, coords, conn, U, dU, stateRef, nthreads=0):
        el, mat = _EL[elType.upper()], _MAT[material.lower()]
        nn = conn.shape[1]
        nd = 3 * nn
        nEl = conn.shape[0]
        props = np.ascontiguousarray(props, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coords, U, dU, stateRef = (np.ascontiguousarray(a, dtype=np.float64) for a in (coords, U, dU, stateRef))
        stateTemp = np.empty_like(stateRef)
        V = np.empty(nEl * nd * nd)
        Pe = np.empty((nEl, nd))
        failed = self.lib.ewo_compute_elements(el, mat, _p(props), nEl, _p(conn), _p(coords), _p(U), _p(dU), _p(stateRef), _p(stateTemp), _p(V), _p(Pe), nthreads)
        return V, Pe, stateTemp, bool(failed)

    def assemble(self, elType, material, props, coords, conn, U, dU, stateRef, x=None, nnz=None, nthreads=0):
        """computeElements + serial P/F scatter + updateCSR (pattern from oracle.port)."""
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        nDof = 3 * coords.shape[0]
        out = {}
        if x is None:
            dofs = port.element_dofs(conn)
            I, J = port.vij_pattern(dofs)  # noqa: E741
            indptr, indices, x = port.csr_pattern(I, J, nDof)
            nnz = indices.size
            out.update(indptr=indptr, indices=indices, x=x)
        V, Pe, stateTemp, failed = self.compute_elements(elType, material, props, coords, conn, U, dU, stateRef, nthreads=nthreads)
        P, F = np.zeros(nDof), np.zeros(nDof)
        self.lib.ewo_scatter_pf(conn.shape[0], conn.shape[1], _p(conn), _p(Pe), _p(P), _p(F))
        data = np.empty(nnz)
        x = np.ascontiguousarray(x, dtype=np.int32)
        self.lib.ewo_update_csr(x.size, _p(x), _p(V), _p(data), nnz)
        out.update(V=V, data=data, P=P, F=F, stateTemp=stateTemp, failed=failed)
        return out

    def make_runner(self, elType, material, props, coords, conn, U, dU, stateRef, nthreads=0):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        dofs = port.element_dofs(conn)
        I, J = port.vij_pattern(dofs)  # noqa: E741
        indptr, indices, x = port.csr_pattern(I, J, 3 * coords.shape[0])
        nnz = indices.size

        def run():
            return self.assemble(elType, material, props, coords, conn, U, dU, stateRef, x=x, nnz=nnz, nthreads=nthreads)

        return run


def load():
    global _inst
    if _inst is None:
        _inst = COracle()
    return _inst
