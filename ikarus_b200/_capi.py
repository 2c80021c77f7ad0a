"""ctypes binding of libikb200.so (include/ikb200.h).  No torch types cross this boundary."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libikb200.so")

IKB_ABI_VERSION = 3
OK, EINVAL, ECUDA, ESTATE, ENOTIMPL, EMATERIAL, ENCCL = 0, -1, -2, -3, -4, -5, -6
STRAIN_LINEAR, STRAIN_GL = 0, 1
MAT_LINEAR, MAT_SVK, MAT_NEOHOOKE, MAT_BLATZKO, MAT_HYPERELASTIC = 0, 1, 2, 3, 4
DEV_NONE, DEV_BLATZKO, DEV_OGDEN_TOTAL, DEV_OGDEN_DEVIATORIC, DEV_INVARIANT_BASED, DEV_ARRUDA_BOYCE, DEV_GENT = range(7)
DBC_RAW, DBC_REDUCED, DBC_FULL = 0, 1, 2
SCALAR, VECTOR, MATRIX = 1, 2, 4


class Desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dim", C.c_int32), ("order", C.c_int32), ("strain", C.c_int32),
                ("material", C.c_int32), ("plane_strain", C.c_int32), ("eas_m", C.c_int32), ("device", C.c_int32),
                ("lam", C.c_double), ("mu", C.c_double), ("n_elem", C.c_int64), ("n_dof", C.c_int64),
                ("reduce_tol", C.c_double), ("eas_function", C.c_int32), ("reserved_", C.c_int32)]


class Hyperelastic(C.Structure):
    """ikb_hyperelastic: a law of the principal-stretch framework (hyperelastic/factory.hh)."""
    _fields_ = [("deviatoric", C.c_int32), ("n", C.c_int32), ("volumetric", C.c_int32), ("reserved_", C.c_int32),
                ("pex", C.c_int32 * 3), ("qex", C.c_int32 * 3), ("par", C.c_double * 3), ("ex", C.c_double * 3),
                ("K", C.c_double), ("beta", C.c_double)]


class TcgInfo(C.Structure):
    """ikb_tcg_info (Eigen::TCGInfo + the model terms TrustRegion needs)."""
    _fields_ = [("delta", C.c_double), ("kappa", C.c_double), ("theta", C.c_double), ("mininner", C.c_int64),
                ("max_iters", C.c_int64), ("tol", C.c_double), ("precond", C.c_int32), ("stop_reason", C.c_int32),
                ("iterations", C.c_int64), ("rel_error", C.c_double), ("eta_norm", C.c_double),
                ("g_dot_eta", C.c_double), ("eta_h_eta", C.c_double)]


PRECOND_IDENTITY, PRECOND_DIAGONAL = 0, 1
EAS_STRAIN, EAS_DISPLACEMENT_GRADIENT, EAS_DISPLACEMENT_GRADIENT_TRANSPOSED = 0, 1, 2

# every symbol include/ikb200.h declares (tests check that the library exports all of them)
SYMBOLS = {
    "ikb_create": [C.POINTER(C.c_void_p), C.POINTER(Desc)],
    "ikb_destroy": [C.c_void_p],
    "ikb_last_error": [C.c_void_p, C.c_char_p, C.c_size_t],
    "ikb_upload_mesh": [C.c_void_p, C.c_void_p, C.c_void_p],
    "ikb_upload_dirichlet": [C.c_void_p, C.c_void_p],
    "ikb_build_pattern": [C.c_void_p],
    "ikb_pattern_nnz": [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "ikb_get_pattern": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "ikb_get_constraints_below": [C.c_void_p, C.c_void_p],
    "ikb_element_linear_indices": [C.c_void_p, C.c_int64, C.c_void_p],
    "ikb_set_solution": [C.c_void_p, C.c_void_p],
    "ikb_set_solution_range": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64],
    "ikb_set_parameter": [C.c_void_p, C.c_double],
    "ikb_set_external_load": [C.c_void_p, C.c_void_p, C.c_int],
    "ikb_assemble": [C.c_void_p, C.c_uint, C.c_int],
    "ikb_invalidate": [C.c_void_p],
    "ikb_get_vector": [C.c_void_p, C.c_int, C.c_void_p],
    "ikb_get_scalar": [C.c_void_p, C.POINTER(C.c_double)],
    "ikb_get_matrix_values": [C.c_void_p, C.c_int, C.c_void_p],
    "ikb_get_dense_matrix": [C.c_void_p, C.c_int, C.c_void_p],
    "ikb_vector_norm": [C.c_void_p, C.c_int, C.POINTER(C.c_double)],
    "ikb_set_hyperelastic": [C.c_void_p, C.POINTER(Hyperelastic)],
    "ikb_eas_update": [C.c_void_p, C.c_void_p],
    "ikb_eas_get_alpha": [C.c_void_p, C.c_void_p],
    "ikb_eas_set_alpha": [C.c_void_p, C.c_void_p],
    "ikb_pcg_solve": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_int),
                      C.POINTER(C.c_double)],
    "ikb_update_solution": [C.c_void_p, C.c_int, C.c_void_p],
    "ikb_get_solution": [C.c_void_p, C.c_void_p],
    "ikb_spmv": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "ikb_calculate_at": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p],
    "ikb_idbc_forces": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "ikb_tcg_solve": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(TcgInfo)],
    "ikb_set_row_ownership": [C.c_void_p, C.c_int64, C.c_int64],
    "ikb_nccl_unique_id": [C.c_void_p],
    "ikb_comm_init": [C.c_void_p, C.c_void_p, C.c_int, C.c_int],
    "ikb_halo_intervals": [C.c_void_p, C.c_void_p, C.c_void_p],
    "ikb_halo_exchange": [C.c_void_p, C.c_char_p],
    "ikb_comm_ipc_export": [C.c_void_p, C.c_void_p],
    "ikb_comm_ipc_import": [C.c_void_p, C.c_void_p, C.c_void_p],
    "ikb_stream": [C.c_void_p, C.POINTER(C.c_void_p)],
    "ikb_sync": [C.c_void_p],
    "ikb_launch_count": [C.c_void_p, C.POINTER(C.c_int64)],
    "ikb_device_ptr": [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)],
    "ikb_time_phase": [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_float)],
}

_lib = None


def load():
    """Load the CUDA library.  There is no CPU fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ikarus_b200.build` (nvcc, sm_100a). "
                "ikarus_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = lib
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
