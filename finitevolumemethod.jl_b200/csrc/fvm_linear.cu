// libfvmcuda: linear templates (assembly, SpMV, Tsit5, Krylov).  Filled in below.
#include "fvm_internal.h"

#define TODO(name) return fvm_fail(h, FVM_ERR_STATE, name ": not implemented yet")
extern "C" int32_t fvm_assemble(fvm_handle h, int32_t, double, const double*, const double*, const double*, const double*,
                                const double*, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_assemble"); }
extern "C" int32_t fvm_get_csr_size(fvm_handle h, int64_t*, int64_t*) { if (!h) return FVM_ERR_ARG; TODO("fvm_get_csr_size"); }
extern "C" int32_t fvm_get_csr(fvm_handle h, int32_t*, int32_t*, double*, double*) { if (!h) return FVM_ERR_ARG; TODO("fvm_get_csr"); }
extern "C" int32_t fvm_spmv(fvm_handle h, const double*, double*, int32_t, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_spmv"); }
extern "C" int32_t fvm_spmv_native(fvm_handle h, const double*, double*, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_spmv_native"); }
extern "C" int32_t fvm_tsit5(fvm_handle h, int32_t, double*, double, double, double, int64_t, const double*, double*, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_tsit5"); }
extern "C" int32_t fvm_krylov(fvm_handle h, int32_t, double*, double, int32_t, int32_t*, double*, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_krylov"); }
extern "C" int32_t fvm_shard_init(fvm_handle h, const void*, int32_t, int32_t) { if (!h) return FVM_ERR_ARG; TODO("fvm_shard_init"); }
extern "C" int32_t fvm_nccl_unique_id(void*) { return FVM_ERR_NCCL; }
