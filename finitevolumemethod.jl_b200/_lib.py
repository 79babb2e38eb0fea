"""ctypes binding of libfvmcuda.so (include/fvmcuda.h).  There is no fallback: if the
library is missing or no CUDA device is present every compute call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfvmcuda.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_NCCL, ERR_IO = range(7)


class FVMCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfvmcuda error %d: %s" % (code, msg))
        self.code = code


class UnsupportedClosureError(FVMCudaError, TypeError):
    """A flux/source/condition function that is not in the compiled device registry."""


_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_bp = C.POINTER(C.c_uint8)
c_lp = C.POINTER(C.c_int64)
H = C.c_void_p

_SIGS = {
    "fvm_create": [c_dp, C.c_int64, c_ip, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(H)],
    "fvm_set_boundary_edges": [H, c_ip, C.c_int64],
    "fvm_set_edge_conditions": [H, C.c_int32, c_bp, c_ip],
    "fvm_set_node_conditions": [H, C.c_int32, c_bp, c_ip],
    "fvm_set_condition_fn": [H, C.c_int32, C.c_int32, C.c_int32, c_dp, C.c_int32],
    "fvm_set_flux": [H, C.c_int32, c_dp, C.c_int32],
    "fvm_set_flux_table": [H, c_dp, c_dp],
    "fvm_set_source": [H, C.c_int32, c_dp, C.c_int32],
    "fvm_set_source_table": [H, c_dp],
    "fvm_finalize": [H, C.c_int32, C.c_int32],
    "fvm_plan_selftest": [c_dp, C.c_int64, c_ip, C.c_int64, C.c_int32, C.c_int32, c_ip, C.c_int64, c_bp, C.c_int32, c_lp],
    "fvm_destroy": [H],
    "fvm_rhs": [H, C.c_double, C.c_void_p, C.c_void_p, C.c_int32],
    "fvm_rhs_native": [H, C.c_double, C.c_void_p, C.c_void_p],
    "fvm_apply_dirichlet": [H, C.c_double, C.c_void_p, C.c_int32],
    "fvm_apply_dirichlet_native": [H, C.c_double, C.c_void_p],
    "fvm_to_native": [H, C.c_void_p, C.c_void_p],
    "fvm_from_native": [H, C.c_void_p, C.c_void_p],
    "fvm_stream_synchronize": [H],
    "fvm_get_stream": [H, C.POINTER(C.c_void_p)],
    "fvm_set_profiling": [H, C.c_int32],
    "fvm_get_profile": [H, c_dp, c_lp],
    "fvm_get_geometry": [H, c_dp, c_dp, c_dp, c_dp, c_dp],
    "fvm_check_recompute_geometry": [H, c_lp],
    "fvm_get_permutation": [H, c_ip, c_ip],
    "fvm_get_stats": [H, c_lp],
    "fvm_assemble": [H, C.c_int32, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int32],
    "fvm_get_csr_size": [H, c_lp, c_lp],
    "fvm_get_csr": [H, c_ip, c_ip, c_dp, c_dp],
    "fvm_spmv": [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32],
    "fvm_spmv_native": [H, C.c_void_p, C.c_void_p, C.c_int32],
    "fvm_tsit5": [H, C.c_int32, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int64, c_dp, C.c_void_p, C.c_int32],
    "fvm_tsit5_adaptive": [H, C.c_int32, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, c_dp,
                           C.c_void_p, C.c_int32, c_lp, c_lp],
    "fvm_jacobian": [H, C.c_double, C.c_void_p, C.c_int32],
    "fvm_get_jacobian_size": [H, c_lp, c_lp],
    "fvm_get_jacobian_csr": [H, c_ip, c_ip, c_dp],
    "fvm_eval_points": [H, C.c_double, C.c_void_p, C.c_int32, C.c_int64, c_ip, c_dp, c_dp, c_dp],
    "fvm_krylov": [H, C.c_int32, C.c_void_p, C.c_double, C.c_int32, c_ip, c_dp, C.c_int32],
    "fvm_newton": [H, C.c_double, C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_int32, c_ip, c_dp, c_dp, c_lp, C.c_int32],
    "fvm_shard_init": [H, C.c_void_p, C.c_int32, C.c_int32],
    "fvm_set_ghost_nodes": [H, c_bp],
    "fvm_set_halo": [H, C.c_int32, c_ip, c_ip, c_ip, c_ip, c_ip],
    "fvm_halo_exchange_native": [H, C.c_void_p],
    "fvm_halo_mode": [H, c_ip, c_ip],
    "fvm_nccl_unique_id": [C.c_void_p],
    "fvm_partition_graph": [C.c_int64, c_ip, C.c_int64, C.c_int32, C.c_int32, c_ip],
    "fvm_partition_edge_cut": [C.c_int64, c_ip, C.c_int64, C.c_int32, c_ip, c_lp],
    "fvm_host_register": [C.c_void_p, C.c_int64],
    "fvm_host_unregister": [C.c_void_p],
    # FVMWIRE containers (host only)
    "fvm_wire_create": [C.c_char_p, C.POINTER(H)],
    "fvm_wire_put": [H, C.c_char_p, C.c_int32, C.c_int32, c_lp, C.c_void_p],
    "fvm_wire_open": [C.c_char_p, C.POINTER(H)],
    "fvm_wire_count": [H, c_ip],
    "fvm_wire_info": [H, C.c_int32, C.c_char_p, c_ip, c_ip, c_lp, c_lp],
    "fvm_wire_find": [H, C.c_char_p, c_ip],
    "fvm_wire_get": [H, C.c_int32, C.c_void_p, C.c_int64],
    "fvm_wire_close": [H],
    "fvm_create_from_wire": [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(H)],
}


def exported_symbols():
    """Every entry point include/fvmcuda.h declares (checked by the CPU test-suite)."""
    return sorted(list(_SIGS) + ["fvm_last_error", "fvm_version", "fvm_wire_last_error", "fvm_wire_crc32"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FVMCudaError(ERR_CUDA, "libfvmcuda.so is not built (%s); run `python finitevolumemethod.jl_b200/build.py`. "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in _SIGS.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int32
        L.fvm_last_error.argtypes = [H]
        L.fvm_last_error.restype = C.c_char_p
        L.fvm_version.argtypes = []
        L.fvm_version.restype = C.c_char_p
        L.fvm_wire_last_error.argtypes = [H]
        L.fvm_wire_last_error.restype = C.c_char_p
        L.fvm_wire_crc32.argtypes = [C.c_void_p, C.c_int64]
        L.fvm_wire_crc32.restype = C.c_uint32
        _lib = L
    return _lib


def check(handle, rc):
    if rc != OK:
        msg = lib().fvm_last_error(handle).decode("utf-8", "replace")
        if rc == ERR_UNSUPPORTED:
            raise UnsupportedClosureError(rc, msg)
        raise FVMCudaError(rc, msg)


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def bp(a):
    return None if a is None else a.ctypes.data_as(c_bp)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
