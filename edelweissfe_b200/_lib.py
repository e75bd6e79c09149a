"""ctypes binding of libedelweiss_b200.so (include/edelweiss_b200.h).

The library is built in-tree by `make -C edelweissfe_b200/csrc` (or __graft_entry__.build()).
There is no CPU fallback: if the shared library is missing, loading raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EWB_LIB_PATH", os.path.join(_HERE, "libedelweiss_b200.so"))  # override: experiments only

# enums of include/edelweiss_b200.h
EWB_C3D8, EWB_C3D20, EWB_C3D8TL, EWB_C3D8R, EWB_C3D8E, EWB_C3D20R = 0, 1, 2, 3, 4, 5
EWB_MAT_LINEARELASTIC, EWB_MAT_VONMISES, EWB_MAT_NEOHOOKE_WA, EWB_MAT_NEOHOOKE_WB, EWB_MAT_NEOHOOKE_WC = 0, 1, 2, 3, 4
EWB_OK, EWB_CUTBACK = 0, 1
EWB_FLAG_ACCUMULATE_PF, EWB_FLAG_FORCE_GENERIC, EWB_FLAG_NO_STIFFNESS, EWB_FLAG_SWEEP_V1, EWB_FLAG_TWO_PHASE = 1, 2, 4, 8, 16

ELEMENT_CODES = {"C3D8": EWB_C3D8, "C3D8N": EWB_C3D8, "C3D20": EWB_C3D20, "C3D20N": EWB_C3D20, "C3D8TL": EWB_C3D8TL, "C3D8NTL": EWB_C3D8TL,
                 "C3D8R": EWB_C3D8R, "C3D8E": EWB_C3D8E, "C3D20R": EWB_C3D20R}
ELEMENT_NODES = {EWB_C3D8: 8, EWB_C3D20: 20, EWB_C3D8TL: 8, EWB_C3D8R: 8, EWB_C3D8E: 8, EWB_C3D20R: 20}
ELEMENT_GAUSS = {EWB_C3D8: 8, EWB_C3D20: 27, EWB_C3D8TL: 8, EWB_C3D8R: 1, EWB_C3D8E: 27, EWB_C3D20R: 8}
MATERIAL_CODES = {
    "linearelastic": EWB_MAT_LINEARELASTIC,
    "vonmises": EWB_MAT_VONMISES,
    "neohookewa": EWB_MAT_NEOHOOKE_WA,
    "neohookewb": EWB_MAT_NEOHOOKE_WB,
    "neohookewc": EWB_MAT_NEOHOOKE_WC,
}
MATERIAL_NSTATE = {EWB_MAT_LINEARELASTIC: 0, EWB_MAT_VONMISES: 1, EWB_MAT_NEOHOOKE_WA: 1, EWB_MAT_NEOHOOKE_WB: 1, EWB_MAT_NEOHOOKE_WC: 1}


class EwbBuffers(C.Structure):
    _fields_ = [
        ("coords", C.c_void_p),
        ("U", C.c_void_p),
        ("dU", C.c_void_p),
        ("state_ref", C.c_void_p),
        ("state_temp", C.c_void_p),
        ("csr_data", C.c_void_p),
        ("P", C.c_void_p),
        ("F", C.c_void_p),
        ("vij", C.c_void_p),
    ]


class EwbError(RuntimeError):
    pass


# every symbol include/edelweiss_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ewb_last_error": (C.c_char_p, []),
    "ewb_version": (C.c_int, []),
    "ewb_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, _P, C.c_int]),
    "ewb_plan_destroy": (None, [_P]),
    "ewb_plan_nnz": (C.c_int64, [_P]),
    "ewb_plan_ndof": (C.c_int64, [_P]),
    "ewb_plan_n_gauss": (C.c_int, [_P]),
    "ewb_plan_n_el_dof": (C.c_int, [_P]),
    "ewb_plan_csr_pattern": (C.c_int, [_P, _P, _P, _P]),
    "ewb_plan_slot_map": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "ewb_plan_set_box": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "ewb_plan_is_box": (C.c_int, [_P]),
    "ewb_plan_set_gather_order": (C.c_int, [_P, _P]),
    "ewb_plan_set_element_order": (C.c_int, [_P, _P]),
    "ewb_debug_stream_schedule": (C.c_int64, [C.c_int, C.c_int64, C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P, _P, C.c_int64]),
    "ewb_assemble": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(EwbBuffers), C.POINTER(C.c_double), C.c_double, C.c_int, _P]),
    "ewb_plan_x_chunks": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, _P, C.c_int]),
    "ewb_assemble_chunks": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(EwbBuffers), C.c_int, C.c_int, C.c_int, _P]),
    "ewb_host_register": (C.c_int, [_P, C.c_int64]),
    "ewb_host_unregister": (C.c_int, [_P]),
    "ewb_poll_status": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "ewb_compute_elements_vij": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(EwbBuffers), _P, C.c_int, _P]),
    "ewb_update_csr": (C.c_int, [_P, _P, _P, _P]),
    "ewb_state_to_soa": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "ewb_state_to_aos": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "ewb_apply_dirichlet_k": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "ewb_apply_dirichlet_r": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "ewb_spmv": (C.c_int, [_P, _P, _P, _P, _P]),
    "ewb_pcg_solve": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), _P]),
    "ewb_surface_pressure": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_double, _P, _P]),
    "ewb_body_force": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P, _P]),
    "ewb_interface_add": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "ewb_plan_set_peer": (C.c_int, [_P, _P, _P, _P]),
    "ewb_plan_status_ptr": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "ewb_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "ewb_peer_free": (C.c_int, [_P]),
    "ewb_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ewb_peer_close": (C.c_int, [_P]),
    "ewb_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EwbError(
            f"{LIB_PATH} not found: build it with `make -C edelweissfe_b200/csrc` (nvcc, sm_100a). "
            "edelweissfe_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise EwbError(f"libedelweiss_b200 error {rc}: {load().ewb_last_error().decode()}")
    return rc
