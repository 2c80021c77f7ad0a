"""ctypes wrapper of oracle/cpu_ref.c (the 'port' CPU baseline).  Test/bench infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libikref.so")
_lib = None
MATERIAL_ID = {"linear": 0, "svk": 1, "neohooke": 2}


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "cpu_ref.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB)
        _lib.ikref_assemble.restype = C.c_int
        _lib.ikref_assemble_opt.restype = C.c_int
        _lib.ikref_element.restype = C.c_int
        _lib.ikref_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def element(dim, material, lam, mu, X, u):
    lib = load()
    nd = (2**dim) * dim
    K = np.zeros((nd, nd))
    R = np.zeros(nd)
    X = np.ascontiguousarray(X, float)
    u = np.ascontiguousarray(u, float)
    rc = lib.ikref_element(C.c_int(dim), C.c_int(MATERIAL_ID[material]), C.c_double(lam), C.c_double(mu), _p(X), _p(u),
                           _p(K), _p(R))
    if rc:
        raise FloatingPointError("material failure")
    return K, R


def assemble(dim, material, lam, mu, corner, edofs, linidx, d, nnz, want_K=True, want_R=True, nthreads=1):
    """Raw K values (Eigen value order) and raw R through the reference's loop structure."""
    lib = load()
    corner = np.ascontiguousarray(corner, float)
    edofs = np.ascontiguousarray(edofs, np.int64)
    d = np.ascontiguousarray(d, float)
    vals = np.zeros(nnz) if want_K else None
    R = np.zeros(d.shape[0]) if want_R else None
    li = np.ascontiguousarray(linidx, np.int64) if want_K else None
    rc = lib.ikref_assemble(C.c_int(dim), C.c_int(MATERIAL_ID[material]), C.c_double(lam), C.c_double(mu),
                            C.c_int64(corner.shape[0]), _p(corner), _p(edofs), _p(li), _p(d), _p(vals),
                            C.c_int64(nnz), _p(R), C.c_int64(d.shape[0]), C.c_int(nthreads))
    if rc:
        raise FloatingPointError(f"{rc} elements failed the material check")
    return vals, R


def assemble_opt(dim, material, lam, mu, corner, edofs, linidx, d, nnz, nthreads=1):
    """The "cpu_opt" baseline: factored NeoHooke math as on the device, one fused K+R sweep, OpenMP."""
    lib = load()
    corner = np.ascontiguousarray(corner, float)
    edofs = np.ascontiguousarray(edofs, np.int64)
    d = np.ascontiguousarray(d, float)
    vals, R = np.zeros(nnz), np.zeros(d.shape[0])
    li = np.ascontiguousarray(linidx, np.int64)
    rc = lib.ikref_assemble_opt(C.c_int(dim), C.c_int(MATERIAL_ID[material]), C.c_double(lam), C.c_double(mu),
                                C.c_int64(corner.shape[0]), _p(corner), _p(edofs), _p(li), _p(d), _p(vals),
                                C.c_int64(nnz), _p(R), C.c_int64(d.shape[0]), C.c_int(nthreads))
    if rc < 0:
        raise NotImplementedError("cpu_opt covers NeoHooke with K and R")
    if rc:
        raise FloatingPointError(f"{rc} Gauss points failed the material check")
    return vals, R


def max_threads():
    return load().ikref_max_threads()
