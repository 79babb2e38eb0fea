// libfvmcuda: device-resident solvers on top of the RHS / SpMV kernels.
//
//   fvm_tsit5   fixed-step Tsit5 (OrdinaryDiffEq tableau) with FSAL and the
//               Dirichlet callback of /root/reference/src/solve.jl:133-165; the state never leaves
//               the device between steps.
//   fvm_krylov  Jacobi-preconditioned CG on the symmetrised system (-V A) x = -V b, or BiCGStab on
//               A x = b.  Dot products: warp-shuffle + fixed-shape block reduction into per-block
//               partials, summed by one block in a fixed order -> deterministic, no atomics; all
//               scalars stay on the device, the host only polls a convergence flag.
#include <cmath>

#include "fvm_device.cuh"

int32_t fvm_launch_spmv(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale);

#define NEED_FINAL(h)                                                                       \
    do {                                                                                    \
        if (!(h)) return FVM_ERR_ARG;                                                       \
        if (!(h)->finalized) return fvm_fail((h), FVM_ERR_STATE, "call fvm_finalize first"); \
        FVM_CUDA(h, cudaSetDevice((h)->device));                                            \
    } while (0)

static int32_t ensure_work(fvm_ctx* h, int count) {
    const size_t n = (size_t)h->N * h->neq;
    for (int i = 0; i < count; ++i)
        if (!h->d_work[i]) {
            int32_t rc = fvm_dev_alloc(h, &h->d_work[i], n);
            if (rc) return rc;
        }
    if (!h->d_red) return fvm_dev_alloc(h, &h->d_red, (size_t)8 * 2048 + 64);
    return FVM_OK;
}

// ---- Tsit5 -----------------------------------------------------------------------------------
static const double TS_C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
static const double TS_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

struct LinComb {
    const double* k[6];
    double c[6];
};

template <int S>
__global__ void __launch_bounds__(256) lincomb_kernel(const int64_t n, double* __restrict__ out, const double* __restrict__ u,
                                                       const double dt, const LinComb lc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = lc.c[0] * lc.k[0][i];
#pragma unroll
        for (int j = 1; j < S; ++j) acc += lc.c[j] * lc.k[j][i];
        out[i] = u[i] + dt * acc;
    }
}

static int32_t launch_lincomb(fvm_ctx* h, int S, int64_t n, double* out, const double* u, double dt, double* const* K,
                              const double* coef) {
    LinComb lc{};
    for (int j = 0; j < S; ++j) {
        lc.k[j] = K[j];
        lc.c[j] = coef[j];
    }
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    switch (S) {
        case 1: lincomb_kernel<1><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
        case 2: lincomb_kernel<2><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
        case 3: lincomb_kernel<3><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
        case 4: lincomb_kernel<4><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
        case 5: lincomb_kernel<5><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
        default: lincomb_kernel<6><<<grid, 256, 0, h->stream>>>(n, out, u, dt, lc); break;
    }
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

// saveat times must be sorted, inside [t0, t1] and (fixed step) on the step grid: an entry the stepper never reaches
// would block every later one and leave its rows of the output unwritten
static int32_t check_saveat(fvm_ctx* h, const char* who, int64_t nsave, const double* tsave, double t0, double t1, double dt_grid) {
    const double eps = 1e-9 * std::max(1.0, std::max(std::fabs(t0), std::fabs(t1)));
    for (int64_t k = 0; k < nsave; ++k) {
        if (!(tsave[k] >= t0 - eps && tsave[k] <= t1 + eps)) return fvm_fail(h, FVM_ERR_ARG, std::string(who) + ": a saveat time lies outside [t0, t1]");
        if (k > 0 && !(tsave[k] >= tsave[k - 1])) return fvm_fail(h, FVM_ERR_ARG, std::string(who) + ": saveat times must be sorted");
        if (dt_grid > 0) {
            const double q = (tsave[k] - t0) / dt_grid;
            if (std::fabs(q - std::round(q)) * dt_grid > 1e-6 * dt_grid + eps)
                return fvm_fail(h, FVM_ERR_ARG, std::string(who) + ": a saveat time is not a multiple of dt after t0 (fixed-step integration)");
        }
    }
    return FVM_OK;
}

extern "C" int32_t fvm_tsit5(fvm_handle h, int32_t use_operator, double* u, double t0, double t1, double dt, int64_t nsave,
                             const double* tsave, double* usave, int32_t on_device) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u, "fvm_tsit5: null state");
    FVM_REQUIRE(h, dt > 0 && t1 >= t0, "fvm_tsit5: need dt > 0 and t1 >= t0");
    FVM_REQUIRE(h, nsave == 0 || (tsave && usave), "fvm_tsit5: save buffers missing");
    if (use_operator && !h->csr.assembled) return fvm_fail(h, FVM_ERR_STATE, "fvm_tsit5: call fvm_assemble first");
    const int64_t nsteps = llround((t1 - t0) / dt);
    FVM_REQUIRE(h, std::fabs(t0 + nsteps * dt - t1) <= 1e-9 * std::max(1.0, std::fabs(t1)),
                "fvm_tsit5: dt must divide the time span (fixed-step integration)");
    int32_t rc = check_saveat(h, "fvm_tsit5", nsave, tsave, t0, t1, dt);
    if (rc) return rc;
    if ((rc = ensure_work(h, 9))) return rc;
    if ((rc = fvm_ensure_state(h))) return rc;
    const int64_t n = h->N * h->neq;
    const size_t bytes = sizeof(double) * n;
    double* U = h->d_work[0];
    double* K[7];
    for (int s = 0; s < 7; ++s) K[s] = h->d_work[1 + s];
    double* TMP = h->d_work[8];

    const double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, U, true))) return rc;

    auto f = [&](double* out, double* x, double t) -> int32_t {
        // sharded: the ghost layer of the stage vector is refreshed (overlapped with independent tiles)
        return use_operator ? fvm_apply_spmv(h, x, out, true, false) : fvm_apply_rhs(h, t, x, out);
    };
    int64_t next_save = 0;
    auto save = [&](double tn) -> int32_t {
        while (next_save < nsave && std::fabs(tsave[next_save] - tn) < 0.5 * dt) {
            double* dst = usave + next_save * n;
            if (on_device) {
                int32_t r = fvm_launch_permute(h, U, dst, false);
                if (r) return r;
            } else {
                int32_t r = fvm_launch_permute(h, U, h->d_io, false);
                if (r) return r;
                FVM_CUDA(h, cudaMemcpyAsync(dst, h->d_io, bytes, cudaMemcpyDeviceToHost, h->stream));
                FVM_CUDA(h, cudaStreamSynchronize(h->stream));
            }
            ++next_save;
        }
        return FVM_OK;
    };
    if ((rc = save(t0))) return rc;
    bool has_callback = false;  // sharded: the same FSAL / callback schedule on every rank
    if ((rc = fvm_global_or(h, !use_operator && h->n_dir > 0, &has_callback))) return rc;
    bool have_k1 = false;
    int parity = 0, graph_parity = 0;  // k1/k7 buffer arrangement; a graph only replays in the one it was captured in
    auto do_step = [&](int64_t step) -> int32_t {
        const double t = t0 + step * dt;
        int32_t r;
        if (!have_k1 && (r = f(K[0], U, t))) return r;
        for (int s = 1; s < 6; ++s) {
            if ((r = launch_lincomb(h, s, n, TMP, U, dt, K, TS_A[s]))) return r;
            if ((r = f(K[s], TMP, t + TS_C[s] * dt))) return r;
        }
        if ((r = launch_lincomb(h, 6, n, U, U, dt, K, TS_A[6]))) return r;
        const double tn = t0 + (step + 1) * dt;
        if (has_callback) {
            // the DiscreteCallback modifies u, so the FSAL value is discarded and k1 is re-evaluated
            if ((r = fvm_launch_dirichlet(h, tn, U))) return r;
            have_k1 = false;
        } else {
            if ((r = f(K[6], U, tn))) return r;
            std::swap(K[0], K[6]);
            parity ^= 1;
            have_k1 = true;
        }
        return FVM_OK;
    };
    // Launch-bound meshes (the README 50x50 config runs ~20 kernels of a few microseconds per step):
    // the steady-state step is captured once into a CUDA graph and replayed.  Two steps per graph
    // when FSAL swaps k1/k7 (the pointer arrangement repeats with period 2).  Not used when a kernel
    // argument changes per step (time-dependent condition functions), when sharded (NCCL in the
    // step) or while the per-kernel profiling events are armed.
    const bool graph_ok = (use_operator || !h->time_dependent) && h->nranks == 1 && !h->halo_ready && !h->profiling && nsteps >= 6 &&
                          !getenv("FVM_NO_GRAPH");
    const int per = has_callback ? 1 : 2;
    cudaGraphExec_t exec = nullptr;
    auto save_due = [&](double tn) { return next_save < nsave && std::fabs(tsave[next_save] - tn) < 0.5 * dt; };
    int64_t step = 0;
    while (step < nsteps) {
        const bool can = graph_ok && step >= 1 && step + per <= nsteps && (per == 1 || !save_due(t0 + (step + 1) * dt)) &&
                         (!exec || parity == graph_parity);
        if (can) {
            if (!exec) {
                graph_parity = parity;
                cudaGraph_t graph = nullptr;
                FVM_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                rc = FVM_OK;
                for (int q = 0; q < per && !rc; ++q) rc = do_step(step + q);
                cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                if (rc) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc;
                }
                FVM_CUDA(h, ce);
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                FVM_CUDA(h, ce);
            }
            cudaError_t ce = cudaGraphLaunch(exec, h->stream);
            if (ce != cudaSuccess) {
                cudaGraphExecDestroy(exec);
                FVM_CUDA(h, ce);
            }
            step += per;
        } else {
            if ((rc = do_step(step))) {
                if (exec) cudaGraphExecDestroy(exec);
                return rc;
            }
            step += 1;
        }
        if ((rc = save(t0 + step * dt))) {
            if (exec) cudaGraphExecDestroy(exec);
            return rc;
        }
    }
    if (exec) {
        FVM_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaGraphExecDestroy(exec);
    }
    if (on_device) {
        if ((rc = fvm_launch_permute(h, U, u, false))) return rc;
    } else {
        if ((rc = fvm_launch_permute(h, U, h->d_io, false))) return rc;
        FVM_CUDA(h, cudaMemcpyAsync(u, h->d_io, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (next_save != nsave) return fvm_fail(h, FVM_ERR_ARG, "fvm_tsit5: not every saveat time was reached");
    return FVM_OK;
}

// ---- deterministic reductions (shared by the adaptive stepper and the Krylov solvers) ----------
#define RED_BLOCKS 1184  // 148 SMs x 8
#define RED_THREADS 256

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ partial) {
    __shared__ double sh[NV][RED_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if (lane == 0) sh[q][warp] = v[q];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double s = lane < RED_THREADS / 32 ? sh[q][lane] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) partial[q * RED_BLOCKS + blockIdx.x] = s;
        }
    }
}

// sums the per-block partials of `nv` slots in a fixed order; one block
__device__ __forceinline__ double final_sum(const double* __restrict__ partial, int slot) {
    __shared__ double sh[RED_THREADS / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < RED_BLOCKS; i += RED_THREADS) s += partial[slot * RED_BLOCKS + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < RED_THREADS / 32; ++w) tot += sh[w];
    return tot;
}

// local sums of the per-block partials -> sc[SC_SUM0 + slot]; sharded runs all-reduce them next
__global__ void __launch_bounds__(RED_THREADS) reduce_partials_kernel(const double* partial, int nslots, double* sc, int guard) {
    if (guard && sc[SC_DONE] != 0.0) {
        if (threadIdx.x < nslots) sc[SC_SUM0 + threadIdx.x] = 0.0;  // keep the collective matched, values unused
        return;
    }
    for (int q = 0; q < nslots; ++q) {
        const double v = final_sum(partial, q);
        if (threadIdx.x == 0) sc[SC_SUM0 + q] = v;
    }
}

#define GRID_STRIDE(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// ---- adaptive Tsit5 --------------------------------------------------------------------------
// Embedded 4th-order error estimate (btilde), OrdinaryDiffEq-style scaled RMS norm and PI step-size
// controller (beta1 = 7/50, beta2 = 2/25, gamma = 0.9, qmin = 0.2, qmax = 10, no change for
// 1 <= q <= 1.2).  saveat times are hit exactly (tstops).  SURVEY.md 8f rank 3 / Appendix C; the
// controller constants live in OrdinaryDiffEq, which is not under /root/reference: parity unpinned.
static const double TS_BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                                0.5823571654525552,      -0.45808210592918697,   0.015151515151515152};

struct ErrComb {
    const double* k[7];
    double c[7];
};

__global__ void __launch_bounds__(RED_THREADS)
    tsit5_err_kernel(const int64_t n, const double* __restrict__ u, const double* __restrict__ unew, const double dt,
                     const ErrComb ec, const double abstol, const double reltol, const uint8_t* __restrict__ kind /* [neq][N] or null */,
                     const int64_t n_nodes, const int neq, double* __restrict__ partial) {
    double v[1] = {0.0};
    GRID_STRIDE(i, n) {
        double e = ec.c[0] * ec.k[0][i];
#pragma unroll
        for (int j = 1; j < 7; ++j) e += ec.c[j] * ec.k[j][i];
        e *= dt;
        const double sc = abstol + fmax(fabs(u[i]), fabs(unew[i])) * reltol;
        const double r = e / sc;
        // sharded: ghost entries belong to another rank's norm (their du is 0 here); the step size must not depend on the partition
        if (!kind || kind[(i % neq) * n_nodes + i / neq] != FVM_NODE_GHOST) v[0] += r * r;
    }
    block_reduce_store<1>(v, partial);
}

__global__ void __launch_bounds__(RED_THREADS) count_owned_kernel(const int64_t n, const uint8_t* __restrict__ kind, const int64_t n_nodes,
                                                                  const int neq, double* __restrict__ partial) {
    double v[1] = {0.0};
    GRID_STRIDE(i, n) v[0] += kind[(i % neq) * n_nodes + i / neq] != FVM_NODE_GHOST ? 1.0 : 0.0;
    block_reduce_store<1>(v, partial);
}

__global__ void __launch_bounds__(RED_THREADS) initdt_kernel(const int64_t n, const double* __restrict__ u, const double* __restrict__ f0,
                                                             const double abstol, const double reltol, const uint8_t* __restrict__ kind,
                                                             const int64_t n_nodes, const int neq, double* __restrict__ partial) {
    double v[2] = {0.0, 0.0};
    GRID_STRIDE(i, n) {
        if (kind && kind[(i % neq) * n_nodes + i / neq] == FVM_NODE_GHOST) continue;
        const double sc = abstol + fabs(u[i]) * reltol;
        v[0] += (u[i] / sc) * (u[i] / sc);
        v[1] += (f0[i] / sc) * (f0[i] / sc);
    }
    block_reduce_store<2>(v, partial);
}

extern "C" int32_t fvm_tsit5_adaptive(fvm_handle h, int32_t use_operator, double* u, double t0, double t1, double abstol,
                                      double reltol, double dt0, int64_t nsave, const double* tsave, double* usave,
                                      int32_t on_device, int64_t* n_accept, int64_t* n_reject) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u && t1 >= t0 && abstol > 0 && reltol > 0, "fvm_tsit5_adaptive: bad arguments");
    FVM_REQUIRE(h, nsave == 0 || (tsave && usave), "fvm_tsit5_adaptive: save buffers missing");
    if (use_operator && !h->csr.assembled) return fvm_fail(h, FVM_ERR_STATE, "fvm_tsit5_adaptive: call fvm_assemble first");
    int32_t rc = check_saveat(h, "fvm_tsit5_adaptive", nsave, tsave, t0, t1, 0.0);
    if (rc) return rc;
    if ((rc = ensure_work(h, 10))) return rc;
    if ((rc = fvm_ensure_state(h))) return rc;
    const int64_t n = h->N * h->neq;
    const size_t bytes = sizeof(double) * n;
    double* U = h->d_work[0];
    double* K[7];
    for (int s = 0; s < 7; ++s) K[s] = h->d_work[1 + s];
    double* TMP = h->d_work[8];
    double* UNEW = h->d_work[9];
    double* partial = h->d_red;
    double* sc = h->d_red + 8 * 2048;
    cudaStream_t st = h->stream;
    const double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, st));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, U, true))) return rc;
    auto f = [&](double* out, double* x, double t) -> int32_t {
        return use_operator ? fvm_apply_spmv(h, x, out, true, false) : fvm_apply_rhs(h, t, x, out);
    };
    double hs[SC_N];
    auto sums = [&](int nslots) -> int32_t {  // partials -> sums (all-reduced when sharded) -> host
        reduce_partials_kernel<<<1, RED_THREADS, 0, st>>>(partial, nslots, sc, 0);
        int32_t r = fvm_allreduce_sum(h, sc + SC_SUM0, nslots);
        if (r) return r;
        FVM_CUDA(h, cudaMemcpyAsync(hs, sc, sizeof(double) * SC_N, cudaMemcpyDeviceToHost, st));
        FVM_CUDA(h, cudaStreamSynchronize(st));
        return FVM_OK;
    };
    int64_t next_save = 0;
    auto save = [&](double tn) -> int32_t {
        while (next_save < nsave && std::fabs(tsave[next_save] - tn) <= 1e-12 * std::max(1.0, std::fabs(tn))) {
            double* dst = usave + next_save * n;
            int32_t r = fvm_launch_permute(h, U, on_device ? dst : h->d_io, false);
            if (r) return r;
            if (!on_device) {
                FVM_CUDA(h, cudaMemcpyAsync(dst, h->d_io, bytes, cudaMemcpyDeviceToHost, st));
                FVM_CUDA(h, cudaStreamSynchronize(st));
            }
            ++next_save;
        }
        return FVM_OK;
    };
    bool has_callback = false;
    if ((rc = fvm_global_or(h, !use_operator && h->n_dir > 0, &has_callback))) return rc;
    // sharded: every rank must see the same error norm -> global sums over OWNED entries and a global owned-entry count
    // (ghost entries are excluded, so accepted / rejected steps do not depend on the number of shards)
    double n_glob = (double)n;
    const uint8_t* ghost_kind = h->halo_ready ? h->dm.kind : nullptr;
    if (ghost_kind) {
        count_owned_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, ghost_kind, h->N, h->neq, partial);
        if ((rc = sums(1))) return rc;
        n_glob = hs[SC_SUM0];
    }
    double t = t0;
    if ((rc = save(t))) return rc;
    if ((rc = f(K[0], U, t))) return rc;
    double dt = dt0;
    if (!(dt > 0)) {  // Hairer's initial step from |u0| and |f(u0)|
        initdt_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, U, K[0], abstol, reltol, ghost_kind, h->N, h->neq, partial);
        if ((rc = sums(2))) return rc;
        const double d0 = std::sqrt(hs[SC_SUM0] / n_glob), d1 = std::sqrt(hs[SC_SUM1] / n_glob);
        dt = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt = std::min(dt, t1 - t0);
    }
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    int64_t nacc = 0, nrej = 0;
    bool have_k1 = true;
    const double t_eps = 1e-12 * std::max(1.0, std::fabs(t1));
    while (t < t1 - t_eps) {
        double tstop = t1;
        if (next_save < nsave && tsave[next_save] > t + t_eps) tstop = std::min(tstop, tsave[next_save]);
        bool clipped = false;
        double dt_try = dt;
        if (t + dt_try >= tstop - t_eps) {
            dt_try = tstop - t;
            clipped = true;
        }
        if (!have_k1 && (rc = f(K[0], U, t))) return rc;
        have_k1 = true;
        for (int s = 1; s < 6; ++s) {
            if ((rc = launch_lincomb(h, s, n, TMP, U, dt_try, K, TS_A[s]))) return rc;
            if ((rc = f(K[s], TMP, t + TS_C[s] * dt_try))) return rc;
        }
        if ((rc = launch_lincomb(h, 6, n, UNEW, U, dt_try, K, TS_A[6]))) return rc;
        if ((rc = f(K[6], UNEW, t + dt_try))) return rc;
        ErrComb ec{};
        for (int j = 0; j < 7; ++j) {
            ec.k[j] = K[j];
            ec.c[j] = TS_BT[j];
        }
        tsit5_err_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, U, UNEW, dt_try, ec, abstol, reltol, ghost_kind, h->N, h->neq, partial);
        if ((rc = sums(1))) return rc;
        const double EEst = std::sqrt(hs[SC_SUM0] / n_glob);
        if (!(EEst == EEst)) return fvm_fail(h, FVM_ERR_ARG, "fvm_tsit5_adaptive: the error estimate is NaN (unstable step)");
        // PI controller
        double q;
        if (EEst == 0.0) q = 1.0 / qmax;
        else {
            const double q11 = std::pow(EEst, beta1);
            q = q11 / std::pow(qold, beta2);
            q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / gamma));
        }
        if (EEst <= 1.0) {
            ++nacc;
            t = clipped ? tstop : t + dt_try;
            std::swap(U, UNEW);
            qold = std::max(EEst, 1e-4);
            double dt_new = dt_try / q;
            if (dt_new >= dt_try && dt_new <= 1.2 * dt_try) dt_new = dt_try;  // qsteady_min = 1, qsteady_max = 1.2
            if (!clipped || dt_new < dt) dt = dt_new;
            if (has_callback) {
                if ((rc = fvm_launch_dirichlet(h, t, U))) return rc;
                have_k1 = false;
            } else {
                std::swap(K[0], K[6]);
            }
            if ((rc = save(t))) return rc;
        } else {
            ++nrej;
            dt = dt_try / std::min(1.0 / qmin, q);
            if (nrej > 100000) return fvm_fail(h, FVM_ERR_ARG, "fvm_tsit5_adaptive: too many rejected steps");
        }
    }
    if ((rc = fvm_launch_permute(h, U, on_device ? u : h->d_io, false))) return rc;
    if (!on_device) FVM_CUDA(h, cudaMemcpyAsync(u, h->d_io, bytes, cudaMemcpyDeviceToHost, st));
    FVM_CUDA(h, cudaStreamSynchronize(st));
    if (n_accept) *n_accept = nacc;
    if (n_reject) *n_reject = nrej;
    if (next_save != nsave) return fvm_fail(h, FVM_ERR_ARG, "fvm_tsit5_adaptive: not every saveat time was reached");
    return FVM_OK;
}

// ---- Krylov ----------------------------------------------------------------------------------

// PCG -------------------------------------------------------------------------------------------
__global__ void pcg_init1_kernel(int64_t n, const double* b, const double* rowscale, double* x) {
    GRID_STRIDE(i, n) if (rowscale[i] == 1.0) x[i] = b[i];  // identity rows: Dirichlet-consistent start
}
__global__ void __launch_bounds__(RED_THREADS)
    pcg_init2_kernel(int64_t n, const double* b, const double* rowscale, const double* Ax_scaled, const double* dinv, double* r,
                     double* z, double* p, double* partial) {
    double v[3] = {0, 0, 0};
    GRID_STRIDE(i, n) {
        const double c = rowscale[i] * b[i];
        const double ri = c - Ax_scaled[i];
        const double zi = dinv[i] * ri;
        r[i] = ri;
        z[i] = zi;
        p[i] = zi;
        const double ru = ri / rowscale[i];  // residual of the unscaled system A x = b
        v[0] += ri * zi;
        v[1] += ru * ru;
        v[2] += b[i] * b[i];
    }
    block_reduce_store<3>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) pcg_init3_kernel(const double* partial, double* sc, double rtol) {
    const double rz = sc[SC_SUM0], rr = sc[SC_SUM1], bb = sc[SC_SUM2];
    if (threadIdx.x == 0) {
        sc[SC_RZ] = rz;
        sc[SC_RR] = rr;
        sc[SC_BNORM2] = bb;
        sc[SC_TOL2] = rtol * rtol * bb;
        sc[SC_ITER] = 0.0;
        sc[SC_DONE] = (rr <= rtol * rtol * bb) ? 1.0 : 0.0;
    }
}
__global__ void __launch_bounds__(RED_THREADS) dot_kernel(int64_t n, const double* a, const double* b, double* partial, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    double v[1] = {0};
    GRID_STRIDE(i, n) v[0] += a[i] * b[i];
    block_reduce_store<1>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) pcg_alpha_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double pq = sc[SC_SUM0];
    if (threadIdx.x == 0) {
        sc[SC_PQ] = pq;
        sc[SC_ALPHA] = sc[SC_RZ] / pq;
    }
}
__global__ void __launch_bounds__(RED_THREADS)
    pcg_update_kernel(int64_t n, const double* p, const double* q, const double* dinv, const double* rowscale, double* x,
                      double* r, double* z, double* partial, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double alpha = sc[SC_ALPHA];
    double v[2] = {0, 0};
    GRID_STRIDE(i, n) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        const double zi = dinv[i] * ri;
        r[i] = ri;
        z[i] = zi;
        const double ru = ri / rowscale[i];
        v[0] += ri * zi;
        v[1] += ru * ru;
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) pcg_beta_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double rz = sc[SC_SUM0], rr = sc[SC_SUM1];
    if (threadIdx.x == 0) {
        sc[SC_BETA] = rz / sc[SC_RZ];
        sc[SC_RZ] = rz;
        sc[SC_RR] = rr;
        sc[SC_ITER] += 1.0;
        if (rr <= sc[SC_TOL2] || !(rr == rr)) sc[SC_DONE] = 1.0;
    }
}
// fused iteration (single GPU, tile kernels): p.q comes out of the SpMV as one partial per CTA; this kernel sums them in a
// fixed order and forms alpha; pcg_beta_fused_kernel does the same for the two sums of the update kernel
__global__ void __launch_bounds__(1024) pcg_alpha_fused_kernel(const double* __restrict__ dotpart, const int count, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += 1024) s += dotpart[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double pq = 0.0;
        for (int w = 0; w < 32; ++w) pq += sh[w];
        sc[SC_PQ] = pq;
        sc[SC_ALPHA] = sc[SC_RZ] / pq;
    }
}
__global__ void __launch_bounds__(RED_THREADS) pcg_beta_fused_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double rz = final_sum(partial, 0);
    __syncthreads();
    const double rr = final_sum(partial, 1);
    if (threadIdx.x == 0) {
        sc[SC_BETA] = rz / sc[SC_RZ];
        sc[SC_RZ] = rz;
        sc[SC_RR] = rr;
        sc[SC_ITER] += 1.0;
        if (rr <= sc[SC_TOL2] || !(rr == rr)) sc[SC_DONE] = 1.0;
    }
}
// Small systems (few SpMV partials): the two 1-block kernels disappear as well -- EVERY block of the update kernel sums the
// SpMV's partials itself (same fixed order in every block, a few KB out of L2) and forms alpha; every block of the
// p-update sums the update kernel's partials and forms beta.  r.z ping-pongs between two scalar slots (`par`), so that
// block 0 can publish the new value while other blocks still read the old one.  4 launches per iteration.
__device__ __forceinline__ double block_sum_fixed(const double* __restrict__ v, const int count) {
    __shared__ double sh[RED_THREADS / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += RED_THREADS) s += __ldcg(v + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();  // sh may still be read by an earlier call
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; ++w) tot += sh[w];
    return tot;
}
__global__ void __launch_bounds__(RED_THREADS)
    pcg_update_merged_kernel(int64_t n, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
                             const double* __restrict__ rowscale, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                             double* __restrict__ partial, double* sc, const double* __restrict__ dotpart, const int count, const int par) {
    if (sc[SC_DONE] != 0.0) return;
    const double pq = block_sum_fixed(dotpart, count);
    const double alpha = sc[par ? SC_RZ2 : SC_RZ] / pq;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc[SC_PQ] = pq;
        sc[SC_ALPHA] = alpha;
    }
    double v[2] = {0, 0};
    GRID_STRIDE(i, n) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        const double zi = dinv[i] * ri;
        r[i] = ri;
        z[i] = zi;
        const double ru = ri / rowscale[i];
        v[0] += ri * zi;
        v[1] += ru * ru;
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS)
    pcg_p_merged_kernel(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ partial, double* sc,
                        const int par) {
    if (sc[SC_DONE] != 0.0) return;  // (set by block 0 of this very launch only when the iteration is over: p is dead then)
    const double rz = block_sum_fixed(partial, RED_BLOCKS);
    const double beta = rz / sc[par ? SC_RZ2 : SC_RZ];
    if (blockIdx.x == 0) {
        const double rr = block_sum_fixed(partial + RED_BLOCKS, RED_BLOCKS);
        if (threadIdx.x == 0) {
            sc[SC_BETA] = beta;
            sc[par ? SC_RZ : SC_RZ2] = rz;
            sc[SC_RR] = rr;
            sc[SC_ITER] += 1.0;
            if (rr <= sc[SC_TOL2] || !(rr == rr)) sc[SC_DONE] = 1.0;
        }
    }
    GRID_STRIDE(i, n) p[i] = z[i] + beta * p[i];
}
__global__ void pcg_p_kernel(int64_t n, const double* z, double* p, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double beta = sc[SC_BETA];
    GRID_STRIDE(i, n) p[i] = z[i] + beta * p[i];
}

// BiCGStab ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
    bicg_init_kernel(int64_t n, const double* b, const double* Ax, double* r, double* rhat, double* p, double* v, double* partial) {
    double s[2] = {0, 0};
    GRID_STRIDE(i, n) {
        const double ri = b[i] - Ax[i];
        r[i] = ri;
        rhat[i] = ri;
        p[i] = 0.0;
        v[i] = 0.0;
        s[0] += ri * ri;
        s[1] += b[i] * b[i];
    }
    block_reduce_store<2>(s, partial);
}
__global__ void __launch_bounds__(RED_THREADS) bicg_init2_kernel(const double* partial, double* sc, double rtol) {
    const double rr = sc[SC_SUM0], bb = sc[SC_SUM1];
    if (threadIdx.x == 0) {
        sc[SC_RR] = rr;
        sc[SC_BNORM2] = bb;
        sc[SC_TOL2] = rtol * rtol * bb;
        sc[SC_RHO] = 1.0;
        sc[SC_ALPHA] = 1.0;
        sc[SC_OMEGA] = 1.0;
        sc[SC_RZ] = rr;  // rhat . r
        sc[SC_ITER] = 0.0;
        sc[SC_RESTART] = 0.0;
        sc[SC_DONE] = (rr <= rtol * rtol * bb) ? 1.0 : 0.0;
    }
}
// beta = (rho_new/rho)(alpha/omega); p = r + beta (p - omega v); y = K^-1 p
__global__ void bicg_p_kernel(int64_t n, const double* r, const double* v, const double* kinv, double* p, double* y, double* rhat,
                              const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    if (sc[SC_RESTART] != 0.0) {  // rhat . r vanished: restart the recurrence with rhat = r
        GRID_STRIDE(i, n) {
            const double ri = r[i];
            rhat[i] = ri;
            p[i] = ri;
            y[i] = kinv[i] * ri;
        }
        return;
    }
    const double beta = (sc[SC_RZ] / sc[SC_RHO]) * (sc[SC_ALPHA] / sc[SC_OMEGA]);
    const double omega = sc[SC_OMEGA];
    GRID_STRIDE(i, n) {
        const double pi = r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        y[i] = kinv[i] * pi;
    }
}
__global__ void __launch_bounds__(RED_THREADS) bicg_alpha_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double rhv = sc[SC_SUM0];
    if (threadIdx.x == 0) {
        sc[SC_RHO] = sc[SC_RZ];
        sc[SC_ALPHA] = sc[SC_RZ] / rhv;
        sc[SC_RESTART] = 0.0;
    }
}
// s = r - alpha v ; z = K^-1 s
__global__ void bicg_s_kernel(int64_t n, const double* r, const double* v, const double* kinv, double* s, double* z, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double alpha = sc[SC_ALPHA];
    GRID_STRIDE(i, n) {
        const double si = r[i] - alpha * v[i];
        s[i] = si;
        z[i] = kinv[i] * si;
    }
}
__global__ void __launch_bounds__(RED_THREADS) bicg_ts_kernel(int64_t n, const double* t, const double* s, double* partial, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    double v[2] = {0, 0};
    GRID_STRIDE(i, n) {
        v[0] += t[i] * s[i];
        v[1] += t[i] * t[i];
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) bicg_omega_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double ts = sc[SC_SUM0], tt = sc[SC_SUM1];
    if (threadIdx.x == 0) sc[SC_OMEGA] = tt != 0.0 ? ts / tt : 0.0;
}
// x += alpha y + omega z ; r = s - omega t ; dots rhat.r, r.r
__global__ void __launch_bounds__(RED_THREADS)
    bicg_x_kernel(int64_t n, const double* y, const double* z, const double* s, const double* t, const double* rhat, double* x,
                  double* r, double* partial, const double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double alpha = sc[SC_ALPHA], omega = sc[SC_OMEGA];
    double v[2] = {0, 0};
    GRID_STRIDE(i, n) {
        x[i] += alpha * y[i] + omega * z[i];
        const double ri = s[i] - omega * t[i];
        r[i] = ri;
        v[0] += rhat[i] * ri;
        v[1] += ri * ri;
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) bicg_end_kernel(const double* partial, double* sc) {
    if (sc[SC_DONE] != 0.0) return;
    const double rhr = sc[SC_SUM0], rr = sc[SC_SUM1];
    if (threadIdx.x == 0) {
        sc[SC_RZ] = rhr;
        sc[SC_RR] = rr;
        sc[SC_ITER] += 1.0;
        if (rr <= sc[SC_TOL2] || !(rr == rr)) {
            sc[SC_DONE] = 1.0;
        } else if (fabs(rhr) <= 1e-28 * rr || sc[SC_OMEGA] == 0.0) {
            sc[SC_RZ] = rr;
            sc[SC_RHO] = sc[SC_ALPHA] = sc[SC_OMEGA] = 1.0;
            sc[SC_RESTART] = 1.0;
        }
    }
}
__global__ void kinv_kernel(int64_t n, const double* dinv, const double* rowscale, double* kinv) {
    GRID_STRIDE(i, n) kinv[i] = dinv[i] * rowscale[i];
}
__global__ void __launch_bounds__(RED_THREADS) resid_kernel(int64_t n, const double* b, const double* Ax, double* partial) {
    double v[2] = {0, 0};
    GRID_STRIDE(i, n) {
        const double ri = b[i] - Ax[i];
        v[0] += ri * ri;
        v[1] += b[i] * b[i];
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(RED_THREADS) resid_final_kernel(const double* partial, double* sc) {
    const double rr = sc[SC_SUM0], bb = sc[SC_SUM1];
    if (threadIdx.x == 0) {
        sc[SC_RR] = rr;
        sc[SC_BNORM2] = bb;
    }
}

extern "C" int32_t fvm_krylov(fvm_handle h, int32_t method, double* x, double rtol, int32_t maxit, int32_t* iters,
                              double* relres, int32_t on_device) {
    NEED_FINAL(h);
    if (!h->csr.assembled) return fvm_fail(h, FVM_ERR_STATE, "fvm_krylov: call fvm_assemble first");
    FVM_REQUIRE(h, x && rtol > 0 && maxit > 0, "fvm_krylov: bad arguments");
    FVM_REQUIRE(h, method == FVM_KRYLOV_PCG || method == FVM_KRYLOV_BICGSTAB, "fvm_krylov: unknown method");
    int32_t rc = ensure_work(h, 10);
    if (rc) return rc;
    if ((rc = fvm_ensure_state(h))) return rc;
    const Csr& c = h->csr;
    const int64_t n = h->N;
    const size_t bytes = sizeof(double) * n;
    double* X = h->d_work[0];
    double* partial = h->d_red;
    double* sc = h->d_red + 8 * 2048;
    cudaStream_t st = h->stream;
    const double* src = x;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, x, bytes, cudaMemcpyHostToDevice, st));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, X, true))) return rc;
    const int G = RED_BLOCKS, B = RED_THREADS;
    // PCG on one GPU with the tile kernels runs the fused iteration (see below)
    const bool fused_pcg = method == FVM_KRYLOV_PCG && !h->halo_ready && h->nranks == 1 && fvm_spmv_fusable(h) && !getenv("FVM_NO_FUSE");
    const bool merged_pcg = fused_pcg && fvm_spmv_fused_partials(h) <= 4096 && !getenv("FVM_PCG_NO_MERGE");
    const int check_every = merged_pcg ? 24 : 25;
    if (fused_pcg) {
        const int32_t np = fvm_spmv_fused_partials(h);
        if (h->dotpart_n < np) {
            if ((rc = fvm_dev_alloc(h, &h->d_dotpart, (size_t)np))) return rc;
            h->dotpart_n = np;
        }
    }
    // local block partials -> local sums -> (sharded) NCCL all-reduce; ghost rows of A, b are zero, so
    // r, z, q, v, t vanish on ghost nodes and the local dot products count every node exactly once
    auto reduce = [&](int nslots, int guard) -> int32_t {
        reduce_partials_kernel<<<1, B, 0, st>>>(partial, nslots, sc, guard);
        return fvm_allreduce_sum(h, sc + SC_SUM0, nslots);
    };
    auto spmv = [&](double* in, double* out, bool scale) -> int32_t {
        return fvm_apply_spmv(h, in, out, false, scale);
    };
    double hsc[SC_N];
    auto poll = [&]() -> int32_t {
        FVM_CUDA(h, cudaMemcpyAsync(hsc, sc, sizeof(double) * SC_N, cudaMemcpyDeviceToHost, st));
        FVM_CUDA(h, cudaStreamSynchronize(st));
        return FVM_OK;
    };
    // The iteration body is a dozen short kernels whose scalars (alpha, beta, omega, the convergence flag)
    // live on the device, so a block of `check_every` iterations is captured once into a CUDA graph and
    // replayed between two polls of the flag: no launch gaps on launch-bound meshes.  Not used when sharded
    // (NCCL inside the iteration) or while the per-kernel profiling events are armed.
    struct Graphs {
        cudaGraphExec_t exec[2] = {nullptr, nullptr};
        ~Graphs() {
            for (cudaGraphExec_t e : exec)
                if (e) cudaGraphExecDestroy(e);
        }
    } graphs;
    const bool graph_ok = h->nranks == 1 && !h->halo_ready && !h->profiling && !getenv("FVM_NO_GRAPH");
    auto run_iterations = [&](auto& iteration, int budget, cudaGraphExec_t& exec) -> int32_t {
        for (int it = 0; it < budget;) {
            const int chunk = std::min(check_every, budget - it);
            if (graph_ok && chunk == check_every) {
                if (!exec) {
                    cudaGraph_t graph = nullptr;
                    FVM_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                    int32_t r = FVM_OK;
                    for (int q = 0; q < chunk && !r; ++q) r = iteration();
                    cudaError_t ce = cudaStreamEndCapture(st, &graph);
                    if (r) {
                        if (graph) cudaGraphDestroy(graph);
                        return r;
                    }
                    FVM_CUDA(h, ce);
                    ce = cudaGraphInstantiate(&exec, graph, 0);
                    cudaGraphDestroy(graph);
                    FVM_CUDA(h, ce);
                }
                FVM_CUDA(h, cudaGraphLaunch(exec, st));
            } else {
                for (int q = 0; q < chunk; ++q) {
                    int32_t r = iteration();
                    if (r) return r;
                }
            }
            it += chunk;
            int32_t r = poll();
            if (r) return r;
            if (hsc[SC_DONE] != 0.0) break;
        }
        return FVM_OK;
    };
    int32_t total_iters = 0;
    double last_rr = -1.0;
    for (int cycle = 0; cycle < 8 && total_iters < maxit; ++cycle) {
        const int budget = maxit - total_iters;
        if (method == FVM_KRYLOV_PCG) {
            double *R = h->d_work[1], *Z = h->d_work[2], *P = h->d_work[3], *Q = h->d_work[4];
            pcg_init1_kernel<<<G, B, 0, st>>>(n, c.b, c.rowscale, X);
            if ((rc = spmv(X, Q, true))) return rc;
            pcg_init2_kernel<<<G, B, 0, st>>>(n, c.b, c.rowscale, Q, c.diag_inv, R, Z, P, partial);
            if ((rc = reduce(3, 0))) return rc;
            pcg_init3_kernel<<<1, B, 0, st>>>(partial, sc, rtol);
            // Fused form (one GPU, tile kernels): p.q comes out of the SpMV's epilogue as one partial per CTA, and the
            // kernels that finish the sums also form alpha / beta: 6 launches and 14 vector passes per iteration instead
            // of 9 and 16
            int par = 0;  // which scalar slot holds r.z (merged form); every cycle starts in slot 0, graphs span an even count
            auto iteration_fused = [&]() -> int32_t {
                SpmvFuse F;
                F.kind = 2;
                F.sc = sc;
                F.dotpart = h->d_dotpart;
                int32_t r;
                if ((r = fvm_apply_spmv_fused(h, P, Q, false, true, F))) return r;
                if (merged_pcg) {
                    pcg_update_merged_kernel<<<G, B, 0, st>>>(n, P, Q, c.diag_inv, c.rowscale, X, R, Z, partial, sc, h->d_dotpart,
                                                              fvm_spmv_fused_partials(h), par);
                    pcg_p_merged_kernel<<<G, B, 0, st>>>(n, Z, P, partial, sc, par);
                    par ^= 1;
                    return FVM_OK;
                }
                pcg_alpha_fused_kernel<<<1, 1024, 0, st>>>(h->d_dotpart, fvm_spmv_fused_partials(h), sc);
                pcg_update_kernel<<<G, B, 0, st>>>(n, P, Q, c.diag_inv, c.rowscale, X, R, Z, partial, sc);
                pcg_beta_fused_kernel<<<1, B, 0, st>>>(partial, sc);
                pcg_p_kernel<<<G, B, 0, st>>>(n, Z, P, sc);
                return FVM_OK;
            };
            auto iteration = [&]() -> int32_t {
                int32_t r;
                if ((r = spmv(P, Q, true))) return r;
                dot_kernel<<<G, B, 0, st>>>(n, P, Q, partial, sc);
                if ((r = reduce(1, 1))) return r;
                pcg_alpha_kernel<<<1, B, 0, st>>>(partial, sc);
                pcg_update_kernel<<<G, B, 0, st>>>(n, P, Q, c.diag_inv, c.rowscale, X, R, Z, partial, sc);
                if ((r = reduce(2, 1))) return r;
                pcg_beta_kernel<<<1, B, 0, st>>>(partial, sc);
                pcg_p_kernel<<<G, B, 0, st>>>(n, Z, P, sc);
                return FVM_OK;
            };
            if (fused_pcg) {
                if ((rc = run_iterations(iteration_fused, budget, graphs.exec[0]))) return rc;
            } else if ((rc = run_iterations(iteration, budget, graphs.exec[0]))) {
                return rc;
            }
        } else {
            double *R = h->d_work[1], *RH = h->d_work[2], *P = h->d_work[3], *V = h->d_work[4], *Y = h->d_work[5],
                   *S = h->d_work[6], *Z = h->d_work[7], *T = h->d_work[8], *KI = h->d_work[9];
            kinv_kernel<<<G, B, 0, st>>>(n, c.diag_inv, c.rowscale, KI);
            pcg_init1_kernel<<<G, B, 0, st>>>(n, c.b, c.rowscale, X);  // identity rows are satisfied from the start
            if ((rc = spmv(X, V, false))) return rc;
            bicg_init_kernel<<<G, B, 0, st>>>(n, c.b, V, R, RH, P, V, partial);
            if ((rc = reduce(2, 0))) return rc;
            bicg_init2_kernel<<<1, B, 0, st>>>(partial, sc, rtol);
            auto iteration = [&]() -> int32_t {
                int32_t r;
                bicg_p_kernel<<<G, B, 0, st>>>(n, R, V, KI, P, Y, RH, sc);
                if ((r = spmv(Y, V, false))) return r;
                dot_kernel<<<G, B, 0, st>>>(n, RH, V, partial, sc);
                if ((r = reduce(1, 1))) return r;
                bicg_alpha_kernel<<<1, B, 0, st>>>(partial, sc);
                bicg_s_kernel<<<G, B, 0, st>>>(n, R, V, KI, S, Z, sc);
                if ((r = spmv(Z, T, false))) return r;
                bicg_ts_kernel<<<G, B, 0, st>>>(n, T, S, partial, sc);
                if ((r = reduce(2, 1))) return r;
                bicg_omega_kernel<<<1, B, 0, st>>>(partial, sc);
                bicg_x_kernel<<<G, B, 0, st>>>(n, Y, Z, S, T, RH, X, R, partial, sc);
                if ((r = reduce(2, 1))) return r;
                bicg_end_kernel<<<1, B, 0, st>>>(partial, sc);
                return FVM_OK;
            };
            if ((rc = run_iterations(iteration, budget, graphs.exec[1]))) return rc;
        }
        FVM_CUDA(h, cudaGetLastError());
        if ((rc = poll())) return rc;
        const int32_t cyc_iters = (int32_t)hsc[SC_ITER];
        total_iters += cyc_iters;
        if (cyc_iters == 0) break;  // the true residual at the start of this cycle already met the tolerance
        if (!(hsc[SC_RR] == hsc[SC_RR])) break;
        if (last_rr >= 0.0 && hsc[SC_RR] >= last_rr && cycle > 1) break;  // no further progress possible
        last_rr = hsc[SC_RR];
    }
    const int32_t done_iters = total_iters;
    // true (unscaled) residual of A x = b
    double* AX = h->d_work[4];
    if ((rc = spmv(X, AX, false))) return rc;
    resid_kernel<<<G, B, 0, st>>>(n, c.b, AX, partial);
    if ((rc = reduce(2, 0))) return rc;
    resid_final_kernel<<<1, B, 0, st>>>(partial, sc);
    if ((rc = poll())) return rc;
    if (iters) *iters = done_iters;
    if (relres) *relres = hsc[SC_BNORM2] > 0 ? std::sqrt(hsc[SC_RR] / hsc[SC_BNORM2]) : std::sqrt(hsc[SC_RR]);
    if (on_device) {
        if ((rc = fvm_launch_permute(h, X, x, false))) return rc;
    } else {
        if ((rc = fvm_launch_permute(h, X, h->d_io, false))) return rc;
        FVM_CUDA(h, cudaMemcpyAsync(x, h->d_io, bytes, cudaMemcpyDeviceToHost, st));
    }
    FVM_CUDA(h, cudaStreamSynchronize(st));
    return FVM_OK;
}


// ---- Newton-Raphson for steady problems, entirely on the device --------------------------------------------
// solve(SteadyFVMProblem(prob), NewtonRaphson()) (/root/reference/src/solve.jl:209-220).  Each iteration evaluates
// F = fvm_eqs!(u) and its sparse Jacobian (fvm_jacobian.cu: dual numbers on the jacobian_sparsity pattern) and solves
// J delta = -F with Jacobi-preconditioned BiCGStab on the block CSR -- nothing but the residual norm crosses PCIe.
// Rows of J without any non-zero (Dirichlet nodes, points that are not vertices: their du is identically 0) act as
// identity rows, so delta stays 0 there and their columns drop out: the system the reference's sparse-direct solver
// sees on the remaining block.

// y = J x on the block CSR (native order, species interleaved); all-zero rows act as identity rows
template <int NEQ>
__global__ void __launch_bounds__(128) jac_spmv_kernel(const int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                        const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double acc[NEQ], mag[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) acc[l] = mag[l] = 0.0;
    for (int e = rowptr[g]; e < rowptr[g + 1]; ++e) {
        const int c = col[e];
#pragma unroll
        for (int l = 0; l < NEQ; ++l)
#pragma unroll
            for (int lp = 0; lp < NEQ; ++lp) {
                const double v = val[((size_t)e * NEQ + l) * NEQ + lp];
                acc[l] += v * x[(size_t)c * NEQ + lp];
                mag[l] += fabs(v);
            }
    }
#pragma unroll
    for (int l = 0; l < NEQ; ++l) y[(size_t)g * NEQ + l] = mag[l] == 0.0 ? x[(size_t)g * NEQ + l] : acc[l];
}
// Jacobi preconditioner 1 / J_ii (1 on identity rows and where the diagonal vanishes); b = -F
template <int NEQ>
__global__ void __launch_bounds__(128) jac_prec_kernel(const int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                        const double* __restrict__ val, const double* __restrict__ F, double* __restrict__ kinv,
                                                        double* __restrict__ b, double* __restrict__ x0) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double d[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) d[l] = 0.0;
    for (int e = rowptr[g]; e < rowptr[g + 1]; ++e)
        if (col[e] == g) {
#pragma unroll
            for (int l = 0; l < NEQ; ++l) d[l] = val[((size_t)e * NEQ + l) * NEQ + l];
        }
#pragma unroll
    for (int l = 0; l < NEQ; ++l) {
        kinv[(size_t)g * NEQ + l] = d[l] != 0.0 ? 1.0 / d[l] : 1.0;
        b[(size_t)g * NEQ + l] = -F[(size_t)g * NEQ + l];
        x0[(size_t)g * NEQ + l] = 0.0;
    }
}
__global__ void __launch_bounds__(RED_THREADS) maxabs_kernel(const int64_t n, const double* __restrict__ v, double* __restrict__ partial) {
    __shared__ double sh[RED_THREADS / 32];
    double m = 0.0;
    GRID_STRIDE(i, n) {
        const double a = fabs(v[i]);
        m = (a > m || a != a) ? a : m;  // NaN propagates
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, m, o);
        m = (w > m || w != w) ? w : m;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RED_THREADS / 32; ++w) m = (sh[w] > m || sh[w] != sh[w]) ? sh[w] : m;
        partial[blockIdx.x] = m;
    }
}
__global__ void axpy1_kernel(const int64_t n, double* __restrict__ u, const double* __restrict__ d) {
    GRID_STRIDE(i, n) u[i] += d[i];
}

extern "C" int32_t fvm_newton(fvm_handle h, double t, double* u, double abstol, double reltol, int32_t maxiters, double lin_rtol,
                              int32_t lin_maxit, int32_t* iters, double* resid, double* resid0, int64_t* lin_iters, int32_t on_device) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u && maxiters >= 0 && lin_rtol > 0 && lin_maxit > 0, "fvm_newton: bad arguments");
    FVM_REQUIRE(h, h->neq <= 2, "fvm_newton: systems with more than 2 species are not compiled");
    if (h->halo_ready) return fvm_fail(h, FVM_ERR_UNSUPPORTED, "fvm_newton: sharded handles are not supported");
    int32_t rc = ensure_work(h, 12);
    if (rc) return rc;
    if ((rc = fvm_ensure_state(h))) return rc;
    const int64_t n = h->N * h->neq;
    const size_t bytes = sizeof(double) * n;
    cudaStream_t st = h->stream;
    double *U = h->d_work[0], *R = h->d_work[1], *RH = h->d_work[2], *P = h->d_work[3], *V = h->d_work[4], *Y = h->d_work[5],
           *S = h->d_work[6], *Z = h->d_work[7], *T = h->d_work[8], *KI = h->d_work[9], *X = h->d_work[10], *F = h->d_work[11];
    double* B = h->d_du;  // right-hand side -F of the Newton system
    double* partial = h->d_red;
    double* sc = h->d_red + 8 * 2048;
    const int G = RED_BLOCKS, Bt = RED_THREADS;
    const double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, st));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, U, true))) return rc;
    if ((rc = fvm_launch_dirichlet(h, t, U))) return rc;  // Dirichlet rows of fvm_eqs! are identically zero: fix the values first
    double hsc[SC_N];
    std::vector<double> hpart(G);
    auto maxabs = [&](const double* v, double* out) -> int32_t {
        maxabs_kernel<<<G, Bt, 0, st>>>(n, v, partial);
        FVM_CUDA(h, cudaMemcpyAsync(hpart.data(), partial, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
        FVM_CUDA(h, cudaStreamSynchronize(st));
        double m = 0.0;
        for (double a : hpart) m = (a > m || a != a) ? a : m;
        *out = m;
        return FVM_OK;
    };
    auto reduce = [&](int nslots, int guard) { reduce_partials_kernel<<<1, Bt, 0, st>>>(partial, nslots, sc, guard); };
    const unsigned jgrid = (unsigned)((h->N + 127) / 128);
    auto J = [&](const double* x, double* y) {
        if (h->neq == 1) jac_spmv_kernel<1><<<jgrid, 128, 0, st>>>((int)h->N, h->csr.rowptr, h->csr.col, h->jac_val, x, y);
        else jac_spmv_kernel<2><<<jgrid, 128, 0, st>>>((int)h->N, h->csr.rowptr, h->csr.col, h->jac_val, x, y);
    };
    if ((rc = fvm_launch_rhs(h, t, U, F))) return rc;
    double res = 0.0, r0 = 0.0;
    if ((rc = maxabs(F, &r0))) return rc;
    res = r0;
    int32_t it = 0;
    int64_t lin_total = 0;
    for (;; ++it) {
        if (res <= abstol + reltol * r0 || it == maxiters || !(res == res) || std::isinf(res)) break;
        if ((rc = fvm_launch_jacobian(h, t, U))) return rc;
        if (h->neq == 1) jac_prec_kernel<1><<<jgrid, 128, 0, st>>>((int)h->N, h->csr.rowptr, h->csr.col, h->jac_val, F, KI, B, X);
        else jac_prec_kernel<2><<<jgrid, 128, 0, st>>>((int)h->N, h->csr.rowptr, h->csr.col, h->jac_val, F, KI, B, X);
        // BiCGStab from x = 0: r = rhat = b
        J(X, V);
        bicg_init_kernel<<<G, Bt, 0, st>>>(n, B, V, R, RH, P, V, partial);
        reduce(2, 0);
        bicg_init2_kernel<<<1, Bt, 0, st>>>(partial, sc, lin_rtol);
        int32_t lit = 0;
        while (lit < lin_maxit) {
            const int chunk = std::min(25, lin_maxit - lit);
            for (int q = 0; q < chunk; ++q) {
                bicg_p_kernel<<<G, Bt, 0, st>>>(n, R, V, KI, P, Y, RH, sc);
                J(Y, V);
                dot_kernel<<<G, Bt, 0, st>>>(n, RH, V, partial, sc);
                reduce(1, 1);
                bicg_alpha_kernel<<<1, Bt, 0, st>>>(partial, sc);
                bicg_s_kernel<<<G, Bt, 0, st>>>(n, R, V, KI, S, Z, sc);
                J(Z, T);
                bicg_ts_kernel<<<G, Bt, 0, st>>>(n, T, S, partial, sc);
                reduce(2, 1);
                bicg_omega_kernel<<<1, Bt, 0, st>>>(partial, sc);
                bicg_x_kernel<<<G, Bt, 0, st>>>(n, Y, Z, S, T, RH, X, R, partial, sc);
                reduce(2, 1);
                bicg_end_kernel<<<1, Bt, 0, st>>>(partial, sc);
            }
            lit += chunk;
            FVM_CUDA(h, cudaGetLastError());
            FVM_CUDA(h, cudaMemcpyAsync(hsc, sc, sizeof(double) * SC_N, cudaMemcpyDeviceToHost, st));
            FVM_CUDA(h, cudaStreamSynchronize(st));
            if (hsc[SC_DONE] != 0.0) break;
        }
        lin_total += (int64_t)hsc[SC_ITER];
        if (!(hsc[SC_RR] == hsc[SC_RR])) return fvm_fail(h, FVM_ERR_ARG, "fvm_newton: the linear solve broke down (NaN residual)");
        axpy1_kernel<<<G, Bt, 0, st>>>(n, U, X);
        if ((rc = fvm_launch_rhs(h, t, U, F))) return rc;
        if ((rc = maxabs(F, &res))) return rc;
    }
    if ((rc = fvm_launch_permute(h, U, on_device ? u : h->d_io, false))) return rc;
    if (!on_device) FVM_CUDA(h, cudaMemcpyAsync(u, h->d_io, bytes, cudaMemcpyDeviceToHost, st));
    FVM_CUDA(h, cudaStreamSynchronize(st));
    if (iters) *iters = it;
    if (resid) *resid = res;
    if (resid0) *resid0 = r0;
    if (lin_iters) *lin_iters = lin_total;
    return FVM_OK;
}
