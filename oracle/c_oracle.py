"""ctypes wrapper of oracle/fvm_oracle_c.c (the timed CPU arm).  TEST INFRASTRUCTURE ONLY:
imported by tests/ and by bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfvmoracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "fvm_oracle_c.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = C.CDLL(_LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_int]
        for f in ("oracle_fvm_eqs_serial", "oracle_fvm_eqs_threaded", "oracle_fvm_eqs_flat"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            getattr(L, f).restype = None
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_volumes.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_spmv.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_spmv.restype = None
        L.oracle_max_threads.restype = C.c_int
        L.oracle_fvm_eqs_general.restype = C.c_int
        L.oracle_fvm_eqs_general.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class COracle:
    """Constant-D diffusion problem with Dirichlet nodes (value 0) on an arbitrary triangle mesh."""

    def __init__(self, points, triangles, dirichlet_nodes, D, nthreads=None):
        L = lib()
        self.xy = np.ascontiguousarray(points, dtype=np.float64)
        self.tri = np.ascontiguousarray(triangles, dtype=np.int32)
        self.dir = np.ascontiguousarray(dirichlet_nodes, dtype=np.int32)
        self.N, self.T = len(self.xy), len(self.tri)
        self.nthreads = int(nthreads or L.oracle_max_threads())
        self.h = L.oracle_create(self.xy.ctypes.data, self.N, self.tri.ctypes.data, self.T, self.dir.ctypes.data,
                                 len(self.dir), float(D), self.nthreads)

    def _call(self, name, u, du=None):
        u = np.ascontiguousarray(u, dtype=np.float64)
        if du is None:
            du = np.empty_like(u)
        getattr(lib(), name)(self.h, u.ctypes.data, du.ctypes.data)
        return du

    def fvm_eqs_serial(self, u, du=None):
        return self._call("oracle_fvm_eqs_serial", u, du)

    def fvm_eqs_threaded(self, u, du=None):
        return self._call("oracle_fvm_eqs_threaded", u, du)

    def fvm_eqs_flat(self, u, du=None):
        return self._call("oracle_fvm_eqs_flat", u, du)

    def volumes(self):
        v = np.empty(self.N)
        lib().oracle_volumes(self.h, v.ctypes.data)
        return v

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmv(rowptr, col, val, b, x, nthreads=1):
    y = np.empty(len(rowptr) - 1)
    lib().oracle_spmv(len(y), rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                      b.ctypes.data if b is not None else None, x.ctypes.data, y.ctypes.data, nthreads)
    return y


def fvm_eqs_general(points, triangles, u, neq=1, flux_model=0, flux_params=(1.0,), src_model=0, src_params=(), dirichlet=None):
    """Serial reference-order RHS for scalar problems / systems with the registered flux and source forms, geometry
    recomputed per triangle (oracle_fvm_eqs_general).  `u`: (N,) or (N, neq); `dirichlet`: bool (neq, N) or (N,) or None.
    Valid where the boundary-edge pass adds nothing (all-Dirichlet boundary or homogeneous Neumann)."""
    xy = np.ascontiguousarray(points, dtype=np.float64)
    tri = np.ascontiguousarray(triangles, dtype=np.int32)
    N = len(xy)
    uu = np.ascontiguousarray(u, dtype=np.float64)
    assert uu.size == N * neq
    du = np.empty_like(uu)
    fp = np.ascontiguousarray(list(flux_params) + [0.0], dtype=np.float64)
    sp = np.ascontiguousarray(list(src_params) + [0.0], dtype=np.float64)
    dk = None
    if dirichlet is not None:
        dk = np.ascontiguousarray(np.broadcast_to(np.asarray(dirichlet, dtype=np.uint8).reshape(-1, N), (neq, N)))
    rc = lib().oracle_fvm_eqs_general(xy.ctypes.data, N, tri.ctypes.data, len(tri), neq, flux_model, fp.ctypes.data, src_model,
                                      sp.ctypes.data, None if dk is None else dk.ctypes.data, uu.ctypes.data, du.ctypes.data)
    if rc:
        raise ValueError("oracle_fvm_eqs_general: bad arguments")
    return du
