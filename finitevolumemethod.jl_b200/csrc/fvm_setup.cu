// libfvmcuda: handle lifecycle and the one-off mesh preprocessing ("finalize").
//
// finalize() replaces the reference's hash containers (Dict{NTuple{3,Int},TriangleProperties},
// Set of triangles, Dicts of conditions; /root/reference/src/geometry.jl:44-49,
// src/conditions.jl:310-316) by flat device arrays laid out for one streaming pass:
//   * triangles are sorted along a Hilbert curve through their centroids and cut into tiles of
//     TT consecutive triangles (a tile is a compact 2-D patch of the mesh);
//   * nodes are renumbered tile-major: a tile's interior nodes (all incident triangles inside
//     the tile) first, then the interface nodes it owns; a tile's nodes are one contiguous range;
//   * every tile gets a node -> (triangle, slot) gather list, so the per-vertex scatter of
//     src/equations/triangle_contributions.jl:10-15 becomes a fixed-order gather in shared
//     memory: deterministic, no fp64 atomics, `du` written exactly once;
//   * interface and boundary nodes are finished by a second small kernel from a buffer of
//     per-tile partial sums whose slots are assigned here.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <parallel/algorithm>
#include <unordered_map>

#include "fvm_internal.h"

static thread_local std::string g_create_err;

// FVM_TIMING=1: wall-clock of the phases of fvm_finalize on stderr
#include <chrono>
#include <omp.h>
struct PhaseTimer {
    bool on = getenv("FVM_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fvm_finalize] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

int32_t fvm_fail(fvm_ctx* h, int32_t code, const std::string& msg) {
    if (h) h->err = msg;
    else g_create_err = msg;
    return code;
}

// Page-locks a caller-owned host buffer (a Julia Vector / NumPy array passed to fvm_rhs, fvm_spmv, ...): the
// host-buffer entry points then copy at full PCIe rate and overlap copy-in, kernels and copy-out; pageable
// memory goes through the driver's staging buffer (measured 19 ms instead of 3.9 ms per fvm_rhs at 16.7M
// nodes).  The caller unregisters before freeing the buffer.
extern "C" int32_t fvm_host_register(void* ptr, int64_t nbytes) {
    if (!ptr || nbytes <= 0) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_host_register: bad arguments");
    cudaError_t e = cudaHostRegister(ptr, (size_t)nbytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return FVM_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();  // do not leave the error behind for an unrelated later call to pick up
        return fvm_fail(nullptr, FVM_ERR_CUDA, std::string("fvm_host_register: ") + cudaGetErrorString(e));
    }
    return FVM_OK;
}

extern "C" int32_t fvm_host_unregister(void* ptr) {
    if (!ptr) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_host_unregister: null pointer");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fvm_fail(nullptr, FVM_ERR_CUDA, std::string("fvm_host_unregister: ") + cudaGetErrorString(e));
    }
    return FVM_OK;
}

extern "C" const char* fvm_version(void) { return "fvmcuda 0.1 (sm_100a)"; }

extern "C" const char* fvm_last_error(fvm_handle h) {
    if (!h) return g_create_err.c_str();
    return h->err.c_str();
}

extern "C" int32_t fvm_create(const double* xy, int64_t N, const int32_t* tri, int64_t T, int32_t index_base,
                              int32_t neq, int32_t device, fvm_handle* out) {
    if (!out) return FVM_ERR_ARG;
    *out = nullptr;
    auto bad = [&](int32_t code, const std::string& m) {
        g_create_err = m;
        return code;
    };
    if (!xy || !tri || N <= 0 || T <= 0) return bad(FVM_ERR_ARG, "fvm_create: empty mesh");
    if (N >= INT32_MAX / 2 || 3 * T >= (int64_t)INT32_MAX * 2)
        return bad(FVM_ERR_ARG, "fvm_create: mesh too large for int32 indices");
    if (neq < 1 || neq > FVM_MAX_NEQ) return bad(FVM_ERR_ARG, "fvm_create: neq must be in 1..4");
    if (index_base != 0 && index_base != 1) return bad(FVM_ERR_ARG, "fvm_create: index_base must be 0 or 1");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return bad(FVM_ERR_CUDA, std::string("fvm_create: no CUDA device (") + cudaGetErrorString(ce) +
                                     "); libfvmcuda has no CPU fallback");
    if (device < 0 || device >= ndev) return bad(FVM_ERR_ARG, "fvm_create: bad device ordinal");
    fvm_ctx* h = new fvm_ctx();
    h->device = device;
    h->neq = neq;
    h->N = N;
    h->T = T;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return bad(FVM_ERR_CUDA, "fvm_create: cannot create stream");
    }
    h->launch_stream = h->stream;
    h->h_xy.assign(xy, xy + 2 * N);
    h->h_tri.resize(3 * T);
    for (int64_t i = 0; i < 3 * T; ++i) {
        int32_t v = tri[i] - index_base;
        if (v < 0 || v >= N) {
            cudaStreamDestroy(h->stream);
            delete h;
            return bad(FVM_ERR_ARG, "fvm_create: triangle vertex out of range");
        }
        h->h_tri[i] = v;
    }
    for (int v = 0; v < neq; ++v) {
        h->h_nkind[v].assign(N, 0);
        h->h_nfidx[v].assign(N, 0);
    }
    h->h_cond.assign((size_t)neq * FVM_MAX_COND_FN, CondFn{FVM_COND_CONST, {0, 0, 0, 0}});
    h->flux.model = FVM_FLUX_DIFF_CONST;
    h->flux.nparams = neq;
    for (int v = 0; v < FVM_MAX_PARAMS; ++v) h->flux.p[v] = 1.0;
    h->source.model = FVM_SRC_ZERO;
    h->h_index_base = index_base;
    *out = h;
    return FVM_OK;
}

extern "C" int32_t fvm_destroy(fvm_handle h) {
    if (!h) return FVM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    fvm_shard_release(h);
    fvm_pipe_release(h);
    for (void* p : h->allocs) cudaFree(p);
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    cudaStreamDestroy(h->stream);
    delete h;
    return FVM_OK;
}

#define NOT_FINAL(h)                                                                      \
    do {                                                                                  \
        if (!(h)) return FVM_ERR_ARG;                                                     \
        if ((h)->finalized) return fvm_fail((h), FVM_ERR_STATE, "mesh already finalized"); \
    } while (0)

extern "C" int32_t fvm_set_boundary_edges(fvm_handle h, const int32_t* uv, int64_t n_edges) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, n_edges >= 0 && (uv || n_edges == 0), "fvm_set_boundary_edges: null edges");
    h->Eb = n_edges;
    h->h_bedge.resize(2 * n_edges);
    for (int64_t i = 0; i < 2 * n_edges; ++i) {
        int32_t v = uv[i] - h->h_index_base;
        FVM_REQUIRE(h, v >= 0 && v < h->N, "fvm_set_boundary_edges: vertex out of range");
        h->h_bedge[i] = v;
    }
    for (int v = 0; v < h->neq; ++v) {
        h->h_ekind[v].assign(n_edges, 0);
        h->h_efidx[v].assign(n_edges, 0);
    }
    return FVM_OK;
}

extern "C" int32_t fvm_set_edge_conditions(fvm_handle h, int32_t var, const uint8_t* kind, const int32_t* fidx) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, var >= 0 && var < h->neq, "fvm_set_edge_conditions: bad species index");
    FVM_REQUIRE(h, kind && fidx, "fvm_set_edge_conditions: null argument");
    for (int64_t e = 0; e < h->Eb; ++e) {
        FVM_REQUIRE(h, kind[e] <= FVM_EDGE_CONSTRAINED, "fvm_set_edge_conditions: bad kind");
        FVM_REQUIRE(h, fidx[e] >= 0 && fidx[e] < FVM_MAX_COND_FN, "fvm_set_edge_conditions: fidx out of range");
    }
    h->h_ekind[var].assign(kind, kind + h->Eb);
    h->h_efidx[var].assign(fidx, fidx + h->Eb);
    return FVM_OK;
}

extern "C" int32_t fvm_set_node_conditions(fvm_handle h, int32_t var, const uint8_t* kind, const int32_t* fidx) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, var >= 0 && var < h->neq, "fvm_set_node_conditions: bad species index");
    FVM_REQUIRE(h, kind && fidx, "fvm_set_node_conditions: null argument");
    for (int64_t i = 0; i < h->N; ++i) {
        FVM_REQUIRE(h, kind[i] <= FVM_NODE_DUDT, "fvm_set_node_conditions: bad kind");
        FVM_REQUIRE(h, fidx[i] >= 0 && fidx[i] < FVM_MAX_COND_FN, "fvm_set_node_conditions: fidx out of range");
    }
    h->h_nkind[var].assign(kind, kind + h->N);
    h->h_nfidx[var].assign(fidx, fidx + h->N);
    return FVM_OK;
}

// Sharded meshes: flags the ghost layer (nodes owned by another rank).  Ghost nodes take part in the
// triangle pass as inputs only; their du is 0 and they are never treated as Dirichlet nodes.
extern "C" int32_t fvm_set_ghost_nodes(fvm_handle h, const uint8_t* is_ghost) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, is_ghost, "fvm_set_ghost_nodes: null argument");
    h->h_ghost.assign(is_ghost, is_ghost + h->N);
    return FVM_OK;
}

extern "C" int32_t fvm_set_condition_fn(fvm_handle h, int32_t var, int32_t fidx, int32_t fn_id, const double* params,
                                        int32_t nparams) {
    if (!h) return FVM_ERR_ARG;
    FVM_REQUIRE(h, var >= 0 && var < h->neq, "fvm_set_condition_fn: bad species index");
    FVM_REQUIRE(h, fidx >= 0 && fidx < FVM_MAX_COND_FN, "fvm_set_condition_fn: fidx out of range");
    if (fn_id < FVM_COND_CONST || fn_id > FVM_COND_EXP_XYT)
        return fvm_fail(h, FVM_ERR_UNSUPPORTED,
                        "fvm_set_condition_fn: condition function is not in the compiled registry "
                        "(arbitrary closures cannot run on the device)");
    static const int need[] = {1, 2, 2, 3, 4};
    FVM_REQUIRE(h, nparams == need[fn_id] && params, "fvm_set_condition_fn: wrong parameter count");
    CondFn c{fn_id, {0, 0, 0, 0}};
    for (int i = 0; i < nparams; ++i) c.p[i] = params[i];
    h->h_cond[(size_t)var * FVM_MAX_COND_FN + fidx] = c;
    h->time_dependent = false;
    for (const CondFn& q : h->h_cond)
        h->time_dependent = h->time_dependent || q.id == FVM_COND_EXP_SAT || (q.id == FVM_COND_EXP_XYT && q.p[3] != 0.0);
    if (h->finalized) {  // condition parameters may change between solves
        FVM_CUDA(h, cudaMemcpyAsync((void*)(h->dm.cond + (size_t)var * FVM_MAX_COND_FN + fidx), &c, sizeof(CondFn),
                                    cudaMemcpyHostToDevice, h->stream));
        FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return FVM_OK;
}

extern "C" int32_t fvm_set_flux(fvm_handle h, int32_t model, const double* params, int32_t nparams) {
    if (!h) return FVM_ERR_ARG;
    const int neq = h->neq;
    int need = -1;
    switch (model) {
        case FVM_FLUX_DIFF_CONST: need = neq; break;
        case FVM_FLUX_DIFF_TABLE: need = 0; break;
        case FVM_FLUX_DIFF_POWER: need = 3 * neq; break;
        case FVM_FLUX_ADVDIFF: need = 3 * neq; break;
        case FVM_FLUX_KELLER_SEGEL: need = 2; break;
        default:
            return fvm_fail(h, FVM_ERR_UNSUPPORTED,
                            "fvm_set_flux: flux model is not in the compiled registry "
                            "(arbitrary closures cannot run on the device)");
    }
    if (model == FVM_FLUX_KELLER_SEGEL && neq != 2)
        return fvm_fail(h, FVM_ERR_ARG, "fvm_set_flux: Keller-Segel needs neq == 2");
    FVM_REQUIRE(h, nparams == need && (params || need == 0), "fvm_set_flux: wrong parameter count");
    if (h->finalized && model == FVM_FLUX_DIFF_TABLE && !h->dm.dtab)
        return fvm_fail(h, FVM_ERR_STATE, "fvm_set_flux: table model needs fvm_set_flux_table before finalize");
    h->flux.model = model;
    h->flux.nparams = nparams;
    for (int i = 0; i < nparams; ++i) h->flux.p[i] = params[i];
    return FVM_OK;
}

extern "C" int32_t fvm_set_flux_table(fvm_handle h, const double* d_cv_edge, const double* d_bnd) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, d_cv_edge, "fvm_set_flux_table: null table");
    FVM_REQUIRE(h, d_bnd || h->Eb == 0, "fvm_set_flux_table: boundary table missing (set boundary edges first)");
    h->h_dtab.assign(d_cv_edge, d_cv_edge + 3 * h->T);
    if (h->Eb) h->h_dbnd.assign(d_bnd, d_bnd + 2 * h->Eb);
    h->flux.model = FVM_FLUX_DIFF_TABLE;
    h->flux.nparams = 0;
    return FVM_OK;
}

extern "C" int32_t fvm_set_source(fvm_handle h, int32_t model, const double* params, int32_t nparams) {
    if (!h) return FVM_ERR_ARG;
    const int neq = h->neq;
    int need = -1;
    switch (model) {
        case FVM_SRC_ZERO: need = 0; break;
        case FVM_SRC_LINEAR: need = 2 * neq; break;
        case FVM_SRC_LOGISTIC: need = neq; break;
        case FVM_SRC_TABLE: need = 0; break;
        case FVM_SRC_GRAY_SCOTT: need = 2; break;
        case FVM_SRC_BRUSSELATOR: need = 0; break;
        case FVM_SRC_KELLER_SEGEL: need = 1; break;
        default:
            return fvm_fail(h, FVM_ERR_UNSUPPORTED,
                            "fvm_set_source: source model is not in the compiled registry "
                            "(arbitrary closures cannot run on the device)");
    }
    if (model >= FVM_SRC_GRAY_SCOTT && neq != 2) return fvm_fail(h, FVM_ERR_ARG, "fvm_set_source: model needs neq == 2");
    FVM_REQUIRE(h, nparams == need && (params || need == 0), "fvm_set_source: wrong parameter count");
    if (model == FVM_SRC_TABLE && h->finalized && !h->dm.src_tab)
        return fvm_fail(h, FVM_ERR_STATE, "fvm_set_source: table model needs fvm_set_source_table before finalize");
    h->source.model = model;
    h->source.nparams = nparams;
    for (int i = 0; i < nparams; ++i) h->source.p[i] = params[i];
    return FVM_OK;
}

extern "C" int32_t fvm_set_source_table(fvm_handle h, const double* s_node) {
    NOT_FINAL(h);
    FVM_REQUIRE(h, s_node, "fvm_set_source_table: null table");
    h->h_srctab.assign(s_node, s_node + h->N * h->neq);
    h->source.model = FVM_SRC_TABLE;
    h->source.nparams = 0;
    return FVM_OK;
}

// Stable parallel LSD radix sort of (key, value) pairs by key, 11 bits per pass; passes whose digit is the same for every
// key are skipped (Hilbert keys of a square domain use 32 of the 52 possible bits).  The result does not depend on the
// number of threads: a stable sort has exactly one outcome.
static void radix_sort_pairs(std::unique_ptr<uint64_t[]>& keys, std::unique_ptr<uint32_t[]>& vals, const int64_t n) {
    if (n < 2) return;
    uint64_t all_or = 0, all_and = ~(uint64_t)0;
#pragma omp parallel for schedule(static) reduction(| : all_or) reduction(& : all_and)
    for (int64_t i = 0; i < n; ++i) {
        all_or |= keys[i];
        all_and &= keys[i];
    }
    const uint64_t varying = all_or ^ all_and;
    constexpr int BITS = 11, BINS = 1 << BITS;
    std::unique_ptr<uint64_t[]> k2(new uint64_t[n]);  // uninitialised: first touched by the threads that fill them
    std::unique_ptr<uint32_t[]> v2(new uint32_t[n]);
    const int nt = omp_get_max_threads();
    std::vector<int64_t> hist((size_t)nt * BINS);
    for (int shift = 0; shift < 64; shift += BITS) {
        if (((varying >> shift) & (BINS - 1)) == 0) continue;
        std::fill(hist.begin(), hist.end(), 0);
        const uint64_t* ks = keys.get();
        const uint32_t* vs = vals.get();
        uint64_t* kd = k2.get();
        uint32_t* vd = v2.get();
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
            int64_t* h = hist.data() + (size_t)t * BINS;
            for (int64_t i = lo; i < hi; ++i) h[(ks[i] >> shift) & (BINS - 1)]++;
#pragma omp barrier
#pragma omp single
            {
                int64_t run = 0;
                for (int d = 0; d < BINS; ++d)
                    for (int q = 0; q < nt; ++q) {
                        const int64_t c = hist[(size_t)q * BINS + d];
                        hist[(size_t)q * BINS + d] = run;
                        run += c;
                    }
            }
            for (int64_t i = lo; i < hi; ++i) {
                const int64_t at = h[(ks[i] >> shift) & (BINS - 1)]++;
                kd[at] = ks[i];
                vd[at] = vs[i];
            }
        }
        keys.swap(k2);
        vals.swap(v2);
    }
}

// ---- Hilbert curve index of a 16-bit lattice point --------------------------------------
static inline uint32_t hilbert_xy2d(uint32_t x, uint32_t y) {
    uint32_t d = 0;
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        uint32_t rx = (x & s) ? 1u : 0u;
        uint32_t ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = 65535u - x;
                y = 65535u - y;
            }
            uint32_t tmp = x;
            x = y;
            y = tmp;
        }
    }
    return d;
}

// Everything fvm_finalize computes on the host before the first device call: the Hilbert tiling, the tile-major
// node numbering, tile-local triangle indices, per-tile gather lists, the interface / partial-slot bookkeeping and
// the live boundary-edge records.  Pure host code on the handle's h_* vectors (fvm_plan_selftest runs it, and
// checks its invariants, without a CUDA device).
struct HostPlan {
    int64_t n_tiles = 0, tpad = 0, n_partial = 0;
    int32_t n_vertices = 0, n_ifc = 0, max_nloc = 0;
    std::vector<int32_t> tile_node0, tile_nint, tile_nown, tile_nloc, tile_ext0, tile_loc0, tile_pp0;
    fvm_rawvec<ushort4> tri_loc;  // (the three T-sized tables are first touched by the tile-parallel passes)
    fvm_rawvec<int32_t> tri_native;
    std::vector<int32_t> ext_ids, ifc_node, ifc_pptr, ppos, live_edges;
    std::vector<uint16_t> inc_ptr;
    fvm_rawvec<uint16_t> inc;
    std::vector<BndEdge> bnd;
    std::vector<double> dbnd_live;
};

static int32_t plan_host(fvm_ctx* h, const int TT, HostPlan& P) {
    const int64_t N = h->N, T = h->T, Eb = h->Eb;
    const int neq = h->neq;
    const double* xy = h->h_xy.data();
    const int32_t* tri = h->h_tri.data();
    PhaseTimer tm;
    // ---- 1. Hilbert sort of triangles by centroid ------------------------------------
    double minx = xy[0], maxx = xy[0], miny = xy[1], maxy = xy[1];
#pragma omp parallel for schedule(static) reduction(min : minx, miny) reduction(max : maxx, maxy)
    for (int64_t i = 0; i < N; ++i) {
        minx = std::min(minx, xy[2 * i]);
        maxx = std::max(maxx, xy[2 * i]);
        miny = std::min(miny, xy[2 * i + 1]);
        maxy = std::max(maxy, xy[2 * i + 1]);
    }
    // The curve runs over near-square BLOCKS laid along the longer side of the bounding box (a row strip of a sharded
    // lattice is 2:1 ... 8:1).  A Hilbert curve enters a square at one corner and leaves at the adjacent one, so the
    // blocks chain without a jump and every tile stays one compact patch; one curve over the bounding square instead
    // leaves and re-enters an elongated domain, and the tiles that span a jump hold 1.6x the nodes (which sizes the
    // shared memory of every CTA).
    const double ex = maxx - minx, ey = maxy - miny;
    const bool long_x = ex >= ey;
    const double e_long = long_x ? ex : ey, e_short = long_x ? ey : ex;
    const int64_t n_blocks = e_short > 0 ? std::max<int64_t>(1, std::min<int64_t>(1 << 20, (int64_t)std::floor(e_long / e_short + 0.5))) : 1;
    const double bw = e_long / (double)n_blocks;
    const double s_long = bw > 0 ? 65535.0 / bw : 0.0, s_short = e_short > 0 ? 65535.0 / e_short : 0.0;
    std::unique_ptr<uint64_t[]> keys(new uint64_t[T]);
    std::unique_ptr<uint32_t[]> order(new uint32_t[T]);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        const int32_t a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        const double cx = (xy[2 * a] + xy[2 * b] + xy[2 * c]) / 3.0, cy = (xy[2 * a + 1] + xy[2 * b + 1] + xy[2 * c + 1]) / 3.0;
        const double l = long_x ? cx - minx : cy - miny, w = long_x ? cy - miny : cx - minx;
        const int64_t blk = bw > 0 ? std::min<int64_t>(n_blocks - 1, std::max<int64_t>(0, (int64_t)(l / bw))) : 0;
        const uint32_t il = (uint32_t)std::min(65535.0, std::max(0.0, (l - (double)blk * bw) * s_long));
        const uint32_t iw = (uint32_t)std::min(65535.0, std::max(0.0, w * s_short));
        // hilbert_xy2d starts at (0, 0) and ends at (65535, 0): its first argument is the direction the blocks advance in
        keys[t] = ((uint64_t)blk << 32) | (uint64_t)hilbert_xy2d(il, iw);
        order[t] = (uint32_t)t;
    }
    radix_sort_pairs(keys, order, T);  // ties keep the caller's triangle order, like the (key, index) pair sort it replaces
    h->tri_old_of_new.resize(T);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) h->tri_old_of_new[t] = (int32_t)order[t];
    keys.reset();
    order.reset();
    const int32_t* told = h->tri_old_of_new.data();
    const int64_t n_tiles = (T + TT - 1) / TT;
    P.n_tiles = n_tiles;
    FVM_REQUIRE(h, n_tiles < INT32_MAX, "too many tiles");

    tm.lap("hilbert sort");
    if (!h->h_ghost.empty())  // ghost nodes are neither free nor Dirichlet
        for (int v = 0; v < neq; ++v)
            for (int64_t i = 0; i < N; ++i)
                if (h->h_ghost[i]) h->h_nkind[v][i] = FVM_NODE_GHOST;
    // ---- 2. boundary edges: adjacent triangle, live flag --------------------------------
    std::vector<uint8_t> forced(N, 0);  // nodes that must be finished by the interface kernel
    std::vector<int32_t>& live_edges = P.live_edges;
    live_edges.clear();
    for (int64_t e = 0; e < Eb; ++e) {
        const int32_t i = h->h_bedge[2 * e], j = h->h_bedge[2 * e + 1];
        bool live = false;
        for (int v = 0; v < neq; ++v) live = live || h->h_nkind[v][i] == FVM_NODE_FREE || h->h_nkind[v][j] == FVM_NODE_FREE;
        if (live) {  // dead work on conditioned edges is skipped
            live_edges.push_back((int32_t)e);
            forced[i] = forced[j] = 1;
        }
    }
    // adjacent triangle (get_adjacent) and stored rotation of EVERY boundary edge; the template
    // assembly needs the dead ones too (abstract_templates.jl:237-267)
    std::vector<int32_t>& all_tri = h->h_edge_tri;
    std::vector<int32_t>& all_rot = h->h_edge_rot;
    all_tri.assign(Eb, -1);
    all_rot.assign(Eb, 0);
    if (Eb > 0) {
        std::unordered_map<uint64_t, int32_t> emap;
        emap.reserve(Eb * 2);
        std::vector<uint8_t> is_src(N, 0);
        for (int64_t e = 0; e < Eb; ++e) {
            emap[((uint64_t)(uint32_t)h->h_bedge[2 * e] << 32) | (uint32_t)h->h_bedge[2 * e + 1]] = (int32_t)e;
            is_src[h->h_bedge[2 * e]] = 1;
        }
        for (int64_t t = 0; t < T; ++t) {
            for (int r = 0; r < 3; ++r) {
                const int32_t a = tri[3 * t + r];
                if (!is_src[a]) continue;
                const int32_t b = tri[3 * t + (r + 1) % 3];
                auto it = emap.find(((uint64_t)(uint32_t)a << 32) | (uint32_t)b);
                if (it != emap.end()) {
                    all_tri[it->second] = (int32_t)t;
                    all_rot[it->second] = r;
                }
            }
        }
        for (int64_t e = 0; e < Eb; ++e)
            FVM_REQUIRE(h, all_tri[e] >= 0, "fvm_finalize: a boundary edge (u,v) is not a ccw edge of any triangle");
    }
    std::vector<int32_t> edge_tri(live_edges.size(), -1), edge_rot(live_edges.size(), 0);
    for (size_t k = 0; k < live_edges.size(); ++k) {
        edge_tri[k] = all_tri[live_edges[k]];
        edge_rot[k] = all_rot[live_edges[k]];
    }

    tm.lap("boundary edges");
    // ---- 3. node classification and tile-major renumbering -------------------------------
    std::vector<int32_t> min_tile(N, INT32_MAX), max_tile(N, -1);
#pragma omp parallel for schedule(static)
    for (int64_t nt = 0; nt < T; ++nt) {  // first / last tile of every vertex: atomic min / max (any order gives the same result)
        const int32_t tile = (int32_t)(nt / TT);
        const int32_t* v = tri + 3 * (int64_t)told[nt];
        for (int r = 0; r < 3; ++r) {
            int32_t cur = __atomic_load_n(&min_tile[v[r]], __ATOMIC_RELAXED);
            while (tile < cur && !__atomic_compare_exchange_n(&min_tile[v[r]], &cur, tile, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            cur = __atomic_load_n(&max_tile[v[r]], __ATOMIC_RELAXED);
            while (tile > cur && !__atomic_compare_exchange_n(&max_tile[v[r]], &cur, tile, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
        }
    }
    std::vector<int32_t>&tile_node0 = P.tile_node0, &tile_nint = P.tile_nint, &tile_nown = P.tile_nown, &tile_nloc = P.tile_nloc;
    tile_node0.assign(n_tiles, 0);
    tile_nint.assign(n_tiles, 0);
    tile_nown.assign(n_tiles, 0);
    tile_nloc.assign(n_tiles, 0);
    h->node_new_of_old.assign(N, -1);
    h->node_old_of_new.assign(N, -1);
    int32_t* new_of_old = h->node_new_of_old.data();
    {
        // a vertex is numbered by the FIRST tile that touches it, so the tiles work independently: pass 1 counts the
        // interior / owned-interface vertices of every tile, a prefix sum gives the tile's first id, pass 2 numbers them in
        // order of first appearance in the tile's triangle list (stamp 1 / 2 marks "seen in pass 1 / 2")
        std::vector<uint8_t> seen(N, 0);
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
            int32_t ni = 0, nf = 0;
            for (int64_t nt = t0; nt < t1; ++nt) {
                const int32_t* v = tri + 3 * (int64_t)told[nt];
                for (int r = 0; r < 3; ++r) {
                    const int32_t g = v[r];
                    if (min_tile[g] != b || seen[g]) continue;
                    seen[g] = 1;
                    if (max_tile[g] == b && !forced[g]) ++ni;
                    else ++nf;
                }
            }
            tile_nint[b] = ni;
            tile_nown[b] = ni + nf;
        }
        int32_t cursor = 0;
        for (int64_t b = 0; b < n_tiles; ++b) {
            tile_node0[b] = cursor;
            cursor += tile_nown[b];
        }
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
            int32_t ci = tile_node0[b], cf = tile_node0[b] + tile_nint[b];
            for (int64_t nt = t0; nt < t1; ++nt) {
                const int32_t* v = tri + 3 * (int64_t)told[nt];
                for (int r = 0; r < 3; ++r) {
                    const int32_t g = v[r];
                    if (min_tile[g] != b || seen[g] == 2) continue;
                    seen[g] = 2;
                    new_of_old[g] = (max_tile[g] == b && !forced[g]) ? ci++ : cf++;
                }
            }
        }
        h->dm.n_vertices = cursor;
        for (int64_t g = 0; g < N; ++g)
            if (new_of_old[g] < 0) new_of_old[g] = cursor++;  // points that are not vertices
    }
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < N; ++g) h->node_old_of_new[new_of_old[g]] = (int32_t)g;
    const int32_t n_vertices = h->dm.n_vertices;
    P.n_vertices = n_vertices;

    tm.lap("node renumbering");
    // ---- 4. tile-local indices, gather lists, interface bookkeeping -----------------------
    const int64_t tpad = n_tiles * TT;
    P.tpad = tpad;
    fvm_rawvec<ushort4>& tri_loc = P.tri_loc;
    tri_loc.resize(tpad);  // every real triangle is written below; the padding of the last tile is zeroed here
    for (int64_t nt = T; nt < tpad; ++nt) tri_loc[nt] = make_ushort4(0, 0, 0, 0);
    fvm_rawvec<int32_t>& tri_native = P.tri_native;  // native node ids per native triangle (geometry kernel input)
    tri_native.resize(3 * T);
    std::vector<int32_t>&tile_ext0 = P.tile_ext0, &tile_loc0 = P.tile_loc0, &tile_pp0 = P.tile_pp0;
    tile_ext0.assign(n_tiles + 1, 0);
    tile_loc0.assign(n_tiles + 1, 0);
    tile_pp0.assign(n_tiles + 1, 0);
    std::vector<int32_t>& ext_ids = P.ext_ids;
    ext_ids.clear();
    std::vector<uint16_t>& inc_ptr = P.inc_ptr;
    fvm_rawvec<uint16_t>& inc = P.inc;
    inc_ptr.clear();
    inc.resize((size_t)3 * tpad);  // 3 entries per real triangle are written below
    for (size_t k = (size_t)3 * TT * (n_tiles - 1) + (size_t)3 * (T - (n_tiles - 1) * TT); k < (size_t)3 * tpad; ++k) inc[k] = 0;
    std::vector<int32_t> ifc_of_new(N, -1);  // compact interface index by native id
    std::vector<int32_t>& ifc_node = P.ifc_node;
    ifc_node.clear();
    for (int64_t b = 0; b < n_tiles; ++b)
        for (int32_t l = tile_nint[b]; l < tile_nown[b]; ++l) {
            ifc_of_new[tile_node0[b] + l] = (int32_t)ifc_node.size();
            ifc_node.push_back(tile_node0[b] + l);
        }
    const int32_t n_ifc = (int32_t)ifc_node.size();
    P.n_ifc = n_ifc;
    std::vector<int32_t> ifc_cnt(n_ifc + 1, 0);
    int32_t& max_nloc = P.max_nloc;
    max_nloc = 0;
    {
        // Tiles are independent: pass A counts every tile's external interface vertices (first-appearance order,
        // deduplicated through a small per-thread hash table), prefix sums place the tile's slices of ext_ids / inc_ptr,
        // pass B fills the tile-local vertex ids, the gather list and the interface counters.
        int hbits = 10;
        while ((1 << hbits) < 8 * TT) ++hbits;
        const int32_t hcap = 1 << hbits, hmask = hcap - 1;
        struct Slot {
            int32_t key, val;
        };
        auto walk_tile = [&](int64_t b, std::vector<Slot>& tab, std::vector<int32_t>& used, bool fill, int32_t* ext_out) -> int32_t {
            const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
            const int32_t node0 = tile_node0[b], nown = tile_nown[b];
            int32_t next = 0;
            for (int64_t nt = t0; nt < t1; ++nt) {
                const int32_t* v = tri + 3 * (int64_t)told[nt];
                uint16_t l3[3];
                for (int r = 0; r < 3; ++r) {
                    const int32_t g = v[r];
                    const int32_t gn = new_of_old[g];
                    int32_t l;
                    if (min_tile[g] == b) l = gn - node0;
                    else {
                        uint32_t q = ((uint32_t)g * 2654435761u) >> (32 - hbits);
                        while (tab[q].key != -1 && tab[q].key != g) q = (q + 1) & hmask;
                        if (tab[q].key == -1) {
                            tab[q].key = g;
                            tab[q].val = nown + next;
                            used.push_back((int32_t)q);
                            if (fill) ext_out[next] = gn;
                            ++next;
                        }
                        l = tab[q].val;
                    }
                    l3[r] = (uint16_t)l;
                    if (fill) tri_native[3 * nt + r] = gn;
                }
                if (fill) tri_loc[nt] = make_ushort4(l3[0], l3[1], l3[2], 0);
            }
            for (int32_t q : used) tab[q].key = -1;
            used.clear();
            return next;
        };
        std::vector<int32_t> next_of(n_tiles, 0);
#pragma omp parallel
        {
            std::vector<Slot> tab(hcap, Slot{-1, 0});
            std::vector<int32_t> used;
#pragma omp for schedule(dynamic, 64)
            for (int64_t b = 0; b < n_tiles; ++b) next_of[b] = walk_tile(b, tab, used, false, nullptr);
        }
        int64_t ext_total = 0, loc_total = 0;
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int32_t nloc = tile_nown[b] + next_of[b];
            FVM_REQUIRE(h, nloc < 65535, "fvm_finalize: tile has too many nodes for 16-bit local ids");
            tile_nloc[b] = nloc;
            max_nloc = std::max(max_nloc, nloc);
            tile_ext0[b] = (int32_t)ext_total;
            ext_total += next_of[b];
            loc_total += loc_total & 1;  // 4-byte aligned per tile (staged with cp.async)
            tile_loc0[b] = (int32_t)loc_total;
            loc_total += nloc + 1;
            FVM_REQUIRE(h, ext_total < INT32_MAX && loc_total < INT32_MAX, "fvm_finalize: tile tables exceed int32");
        }
        tile_ext0[n_tiles] = (int32_t)ext_total;
        tile_loc0[n_tiles] = (int32_t)loc_total;
        ext_ids.assign((size_t)ext_total, 0);
        inc_ptr.assign((size_t)loc_total + 2, 0);  // + slack for the 4-byte granular copy of the last tile
#pragma omp parallel
        {
            std::vector<Slot> tab(hcap, Slot{-1, 0});
            std::vector<int32_t> used, cnt, fill;
#pragma omp for schedule(dynamic, 64)
            for (int64_t b = 0; b < n_tiles; ++b) {
                const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
                const int32_t node0 = tile_node0[b], nown = tile_nown[b], nloc = tile_nloc[b];
                const int32_t next = walk_tile(b, tab, used, true, ext_ids.data() + tile_ext0[b]);
                // gather list: local node -> (local triangle, slot), ascending triangle order
                cnt.assign(nloc + 1, 0);
                for (int64_t nt = t0; nt < t1; ++nt) {
                    const ushort4 q = tri_loc[nt];
                    cnt[q.x + 1]++;
                    cnt[q.y + 1]++;
                    cnt[q.z + 1]++;
                }
                for (int32_t l = 0; l < nloc; ++l) cnt[l + 1] += cnt[l];
                uint16_t* ip = inc_ptr.data() + tile_loc0[b];
                for (int32_t l = 0; l <= nloc; ++l) ip[l] = (uint16_t)cnt[l];
                fill.assign(cnt.begin(), cnt.end() - 1);
                uint16_t* tinc = inc.data() + (size_t)3 * TT * b;
                for (int64_t nt = t0; nt < t1; ++nt) {
                    const ushort4 q = tri_loc[nt];
                    const uint16_t lt = (uint16_t)(nt - t0);
                    tinc[fill[q.x]++] = (uint16_t)(lt << 2 | 0);
                    tinc[fill[q.y]++] = (uint16_t)(lt << 2 | 1);
                    tinc[fill[q.z]++] = (uint16_t)(lt << 2 | 2);
                }
                // interface locals of this tile contribute one partial each
                for (int32_t l = tile_nint[b]; l < nown; ++l) __atomic_fetch_add(&ifc_cnt[ifc_of_new[node0 + l]], 1, __ATOMIC_RELAXED);
                for (int32_t k = 0; k < next; ++k) __atomic_fetch_add(&ifc_cnt[ifc_of_new[ext_ids[tile_ext0[b] + k]]], 1, __ATOMIC_RELAXED);
            }
        }
    }
    for (size_t k = 0; k < live_edges.size(); ++k) {
        const int32_t e = live_edges[k];
        ifc_cnt[ifc_of_new[new_of_old[h->h_bedge[2 * e]]]]++;
        ifc_cnt[ifc_of_new[new_of_old[h->h_bedge[2 * e + 1]]]]++;
    }
    std::vector<int32_t>& ifc_pptr = P.ifc_pptr;
    ifc_pptr.assign(n_ifc + 1, 0);
    for (int32_t k = 0; k < n_ifc; ++k) ifc_pptr[k + 1] = ifc_pptr[k] + ifc_cnt[k];
    const int64_t n_partial = ifc_pptr[n_ifc];
    P.n_partial = n_partial;
    std::vector<int32_t> pfill(ifc_pptr.begin(), ifc_pptr.end() - 1);
    std::vector<int32_t>& ppos = P.ppos;
    ppos.clear();
    for (int64_t b = 0; b < n_tiles; ++b) {  // ascending tile order = fixed summation order
        tile_pp0[b] = (int32_t)ppos.size();
        for (int32_t l = tile_nint[b]; l < tile_nown[b]; ++l) ppos.push_back(pfill[ifc_of_new[tile_node0[b] + l]]++);
        for (int32_t k = tile_ext0[b]; k < tile_ext0[b + 1]; ++k) ppos.push_back(pfill[ifc_of_new[ext_ids[k]]]++);
    }
    tile_pp0[n_tiles] = (int32_t)ppos.size();

    tm.lap("tile lists");
    // ---- 5. live boundary-edge records ------------------------------------------------------
    std::vector<BndEdge>& bnd = P.bnd;
    bnd.assign(live_edges.size(), BndEdge{});
    std::vector<double>& dbnd_live = P.dbnd_live;
    dbnd_live.clear();
    for (size_t k = 0; k < live_edges.size(); ++k) {
        const int32_t e = live_edges[k];
        BndEdge& r = bnd[k];
        const int32_t i = h->h_bedge[2 * e], j = h->h_bedge[2 * e + 1];
        const int32_t* v = tri + 3 * (int64_t)edge_tri[k];
        for (int q = 0; q < 3; ++q) r.v[q] = new_of_old[v[q]];
        r.pi = edge_rot[k];
        r.pj = (edge_rot[k] + 1) % 3;
        r.slot_i = pfill[ifc_of_new[new_of_old[i]]]++;
        r.slot_j = pfill[ifc_of_new[new_of_old[j]]]++;
        r.orig = e;
        r.px = xy[2 * i];
        r.py = xy[2 * i + 1];
        r.qx = xy[2 * j];
        r.qy = xy[2 * j + 1];
        for (int s = 0; s < FVM_MAX_NEQ; ++s) {
            r.kind[s] = s < neq ? h->h_ekind[s][e] : 0;
            r.fidx[s] = s < neq ? h->h_efidx[s][e] : 0;
        }
        if (!h->h_dbnd.empty()) {
            dbnd_live.push_back(h->h_dbnd[2 * e]);
            dbnd_live.push_back(h->h_dbnd[2 * e + 1]);
        }
    }
    return FVM_OK;
}

// ---- tile packs of the streaming recompute kernel (fvm_rhs_stream.cu) ---------------------------------------------
// One 16-byte aligned record per tile (TilePackHdr + sections), built tile-parallel on the host.  The node gather
// list is re-laid out in UNITS of 32 consecutive local nodes (one warp of the node pass): unit q owns
// urow[q+1] - urow[q] rows of 32 uint16 codes, entry j of node 32 q + i at row urow[q] + j, column i.
// Each code is the BYTE offset of a vertex
// contribution inside the tile's contribution plane; padded entries point at a zero word behind the plane.
// Per-node order = the CSR list (ascending triangle).
//
// Bank-conflict-free by construction: the contribution of (triangle lt, slot s) lives in the 16-double block
// s * TT + (lt & ~15) at position color(lt, s).  A half-warp of the triangle pass (16 consecutive lt, one slot) writes one
// block; a half-warp of the node pass (half a unit, one gather row) reads 16 contributions.  Taking the
// write groups and the read groups as the two sides of a bipartite multigraph (one edge per contribution, degree <= 16),
// a proper edge colouring with 16 colours (Koenig) puts every group's 16 accesses in 16 different 8-byte bank pairs:
// neither the scatter nor the gather has a shared-memory bank conflict.  The 4-bit colours travel in the spare 16 bits
// of the triangle's ushort4 record.
static inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }

static inline int32_t unit_rows(const uint16_t* iptr, int32_t nloc, int32_t q) {
    int32_t w = 0;
    for (int32_t l = 32 * q; l < std::min(nloc, 32 * q + 32); ++l) w = std::max<int32_t>(w, iptr[l + 1] - iptr[l]);
    return w;
}

static int32_t build_tile_packs(fvm_ctx* h, const HostPlan& P, const int TT, const double* xyn, const uint8_t* kn,
                                const double* dtab_native /* [3][tpad] or null */, fvm_rawvec<uint8_t>& packs,
                                std::vector<int4>& dir, int32_t& pack_cap) {
    const int64_t n_tiles = P.n_tiles, N = h->N, T = h->T;
    const int neq = h->neq;
    std::vector<int64_t> off(n_tiles + 1, 0);
    std::vector<int32_t> nrows(n_tiles, 0);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < n_tiles; ++b) {
        const int32_t nint = P.tile_nint[b], nloc = P.tile_nloc[b];
        const int32_t ntri = (int32_t)std::min<int64_t>(TT, T - b * TT);
        const uint16_t* iptr = P.inc_ptr.data() + P.tile_loc0[b];
        const int32_t nunit = (nloc + 31) / 32;
        int32_t rows = 0;
        for (int32_t q = 0; q < nunit; ++q) rows += unit_rows(iptr, nloc, q);
        nrows[b] = rows;
        int64_t sz = sizeof(TilePackHdr);
        sz = align16(sz + 8 * (int64_t)ntri);                  // tri: ushort4
        sz = align16(sz + 16 * (int64_t)nloc);                 // xy: double2
        sz = align16(sz + 8 * (int64_t)nint);                  // 1 / V
        sz = align16(sz + (int64_t)neq * align16(nint));       // kind: [neq][align16(nint)]
        sz = align16(sz + 4 * (int64_t)(nloc - nint));         // partial slots
        sz = align16(sz + 2 * (int64_t)(nunit + 1));           // unit row offsets
        sz = align16(sz + 64 * (int64_t)rows);                 // gather rows
        if (dtab_native) sz = align16(sz + 24 * (int64_t)((ntri + 1) & ~1));
        off[b + 1] = sz;
    }
    pack_cap = 0;
    for (int64_t b = 0; b < n_tiles; ++b) {
        pack_cap = std::max<int32_t>(pack_cap, (int32_t)off[b + 1]);
        FVM_REQUIRE(h, nrows[b] < 65535, "fvm_finalize: tile gather list too long for the streaming kernel");
        off[b + 1] += off[b];
    }
    FVM_REQUIRE(h, (off[n_tiles] >> 4) < (int64_t)UINT32_MAX, "fvm_finalize: tile packs exceed 64 GB");
    packs.resize((size_t)off[n_tiles]);
    dir.resize(2 * n_tiles);
#pragma omp parallel
    {
        // scratch of the edge colouring, reused from tile to tile
        std::vector<int16_t> colL, colR;  // [vertex][colour] -> edge using that colour there (or -1)
        std::vector<int32_t> eL, eR, path;
        std::vector<int8_t> ecol;
        std::vector<uint16_t> usedL, usedR;
#pragma omp for schedule(dynamic, 64)
        for (int64_t b = 0; b < n_tiles; ++b) {
            uint8_t* base = packs.data() + off[b];
            std::memset(base, 0, (size_t)(off[b + 1] - off[b]));
            TilePackHdr H{};
            H.node0 = P.tile_node0[b];
            H.nint = P.tile_nint[b];
            H.nown = P.tile_nown[b];
            H.nloc = P.tile_nloc[b];
            H.ntri = (int32_t)std::min<int64_t>(TT, T - b * TT);
            H.nunit = (H.nloc + 31) / 32;
            H.bytes = (int32_t)(off[b + 1] - off[b]);
            int64_t o = sizeof(TilePackHdr);
            H.off_tri = (int32_t)o;
            o = align16(o + 8 * (int64_t)H.ntri);
            H.off_xy = (int32_t)o;
            o = align16(o + 16 * (int64_t)H.nloc);
            H.off_vinv = (int32_t)o;
            o = align16(o + 8 * (int64_t)H.nint);
            H.off_kind = (int32_t)o;
            const int32_t kstride = (int32_t)align16(H.nint);
            o = align16(o + (int64_t)neq * kstride);
            H.off_ppos = (int32_t)o;
            o = align16(o + 4 * (int64_t)(H.nloc - H.nint));
            H.off_urow = (int32_t)o;
            o = align16(o + 2 * (int64_t)(H.nunit + 1));
            H.off_list = (int32_t)o;
            o = align16(o + 64 * (int64_t)nrows[b]);
            H.off_dtab = dtab_native ? (int32_t)o : 0;
            std::memcpy(base, &H, sizeof(H));
            ushort4* tri_out = reinterpret_cast<ushort4*>(base + H.off_tri);
            std::memcpy(tri_out, P.tri_loc.data() + b * TT, 8 * (size_t)H.ntri);
            double* xy = reinterpret_cast<double*>(base + H.off_xy);
            for (int32_t l = 0; l < H.nloc; ++l) {
                const int32_t g = l < H.nown ? H.node0 + l : P.ext_ids[P.tile_ext0[b] + (l - H.nown)];
                xy[2 * l] = xyn[2 * (size_t)g];
                xy[2 * l + 1] = xyn[2 * (size_t)g + 1];
            }
            // 1/V is written on the device once the control volumes exist (fvm_stream_fill_vinv)
            bool all_free = true;
            for (int v = 0; v < neq; ++v) {
                std::memcpy(base + H.off_kind + (size_t)v * kstride, kn + (size_t)v * N + H.node0, (size_t)H.nint);
                for (int32_t l = 0; l < H.nint; ++l) all_free = all_free && kn[(size_t)v * N + H.node0 + l] == FVM_NODE_FREE;
            }
            reinterpret_cast<TilePackHdr*>(base)->flags = all_free ? 1 : 0;
            std::memcpy(base + H.off_ppos, P.ppos.data() + P.tile_pp0[b], 4 * (size_t)(H.nloc - H.nint));
            uint16_t* urow = reinterpret_cast<uint16_t*>(base + H.off_urow);
            uint16_t* list = reinterpret_cast<uint16_t*>(base + H.off_list);
            const uint16_t* iptr = P.inc_ptr.data() + P.tile_loc0[b];
            const uint16_t* inc = P.inc.data() + (size_t)3 * TT * b;
            int32_t row = 0;
            for (int32_t q = 0; q < H.nunit; ++q) {
                urow[q] = (uint16_t)row;
                row += unit_rows(iptr, H.nloc, q);
            }
            urow[H.nunit] = (uint16_t)row;
            // ---- bipartite edge colouring: left = write groups (slot, lt / 16), right = read groups (gather rows) ----
            const int32_t nG = TT / 16, nL = 3 * nG, nR = 2 * row, nE = 3 * H.ntri;
            colL.assign((size_t)nL * 16, -1);
            colR.assign((size_t)std::max(nR, 1) * 16, -1);
            eL.resize(nE);
            eR.resize(nE);
            ecol.assign(nE, -1);
            for (int32_t l = 0; l < H.nloc; ++l)
                for (int32_t q = iptr[l]; q < iptr[l + 1]; ++q) {
                    const uint16_t c = inc[q];
                    const int32_t lt = c >> 2, slot = c & 3, e = 3 * lt + slot;
                    eL[e] = slot * nG + (lt >> 4);
                    eR[e] = 2 * (urow[l >> 5] + (q - iptr[l])) + ((l >> 4) & 1);
                }
            usedL.assign(nL, 0);  // bit c set: colour c is taken at that vertex (same information as colL / colR >= 0)
            usedR.assign(std::max(nR, 1), 0);
            for (int32_t e = 0; e < nE; ++e) {
                int16_t* cl = colL.data() + (size_t)eL[e] * 16;
                int16_t* cr = colR.data() + (size_t)eR[e] * 16;
                const uint32_t fl = (uint16_t)~usedL[eL[e]], fr = (uint16_t)~usedR[eR[e]];  // free colours (a vertex has <= 16 edges)
                int both = (fl & fr) ? __builtin_ctz(fl & fr) : -1;
                const int ca = fl ? __builtin_ctz(fl) : -1, cbb = fr ? __builtin_ctz(fr) : -1;
                if (both < 0) {  // ca free on the left, cbb free on the right: flip the ca/cbb path that starts at the right end
                    path.clear();
                    int32_t cur = eR[e];
                    bool right = true;
                    int want = ca;
                    for (;;) {
                        const int16_t e1 = right ? colR[(size_t)cur * 16 + want] : colL[(size_t)cur * 16 + want];
                        if (e1 < 0) break;
                        path.push_back(e1);
                        cur = right ? eL[e1] : eR[e1];
                        right = !right;
                        want = want == ca ? cbb : ca;
                    }
                    for (int32_t e1 : path) {
                        colL[(size_t)eL[e1] * 16 + ecol[e1]] = -1;
                        colR[(size_t)eR[e1] * 16 + ecol[e1]] = -1;
                        usedL[eL[e1]] &= (uint16_t)~(1u << ecol[e1]);
                        usedR[eR[e1]] &= (uint16_t)~(1u << ecol[e1]);
                    }
                    for (int32_t e1 : path) {
                        ecol[e1] = (int8_t)(ecol[e1] == ca ? cbb : ca);
                        colL[(size_t)eL[e1] * 16 + ecol[e1]] = (int16_t)e1;
                        colR[(size_t)eR[e1] * 16 + ecol[e1]] = (int16_t)e1;
                        usedL[eL[e1]] |= (uint16_t)(1u << ecol[e1]);
                        usedR[eR[e1]] |= (uint16_t)(1u << ecol[e1]);
                    }
                    both = ca;
                }
                ecol[e] = (int8_t)both;
                cl[both] = (int16_t)e;
                cr[both] = (int16_t)e;
                usedL[eL[e]] |= (uint16_t)(1u << both);
                usedR[eR[e]] |= (uint16_t)(1u << both);
            }
            for (int32_t lt = 0; lt < H.ntri; ++lt)
                tri_out[lt].w = (uint16_t)(ecol[3 * lt] | (ecol[3 * lt + 1] << 4) | (ecol[3 * lt + 2] << 8));
            for (int32_t q = 0; q < H.nunit; ++q)
                for (int32_t r = urow[q]; r < urow[q + 1]; ++r)
                    for (int32_t half = 0; half < 2; ++half) {
                        // padded entries of this read group share a zero word in a bank pair no real entry of the group uses
                        const uint32_t fr = (uint16_t)~usedR[2 * r + half];
                        const int freec = fr ? __builtin_ctz(fr) : 0;
                        const int32_t j = r - urow[q];
                        for (int32_t i = 16 * half; i < 16 * half + 16; ++i) {
                            const int32_t l = 32 * q + i;
                            uint16_t code = (uint16_t)((3 * TT + freec) * 8);
                            if (l < H.nloc && iptr[l] + j < iptr[l + 1]) {
                                const uint16_t c = inc[iptr[l] + j];
                                const int32_t lt = c >> 2, slot = c & 3;
                                code = (uint16_t)((slot * TT + (lt & ~15) + ecol[3 * lt + slot]) * 8);
                            }
                            list[(size_t)r * 32 + i] = code;
                        }
                    }
            if (dtab_native) {
                double* dt = reinterpret_cast<double*>(base + H.off_dtab);
                const int32_t np = (H.ntri + 1) & ~1;
                for (int e = 0; e < 3; ++e)
                    for (int32_t lt = 0; lt < H.ntri; ++lt) dt[(size_t)e * np + lt] = dtab_native[(size_t)e * P.tpad + b * TT + lt];
            }
            dir[2 * b] = make_int4((int32_t)(uint32_t)(off[b] >> 4), H.bytes, H.node0, H.nown);
            dir[2 * b + 1] = make_int4(P.tile_ext0[b], H.nloc - H.nown, 0, 0);
        }
    }
    return FVM_OK;
}

// Runs the host planning of fvm_finalize on a mesh WITHOUT a CUDA device and checks the invariants the kernels
// rely on (tests/test_host_cpu.py).  stats: n_tiles, n_vertices, n_interface, n_partial, n_external, max_local_nodes,
// n_live_boundary_edges, gather-list entries.
extern "C" int32_t fvm_plan_selftest(const double* xy, int64_t N, const int32_t* tri, int64_t T, int32_t index_base, int32_t neq,
                                     const int32_t* bedges, int64_t Eb, const uint8_t* node_kind, int32_t tile_triangles,
                                     int64_t* stats) {
    if (!xy || !tri || N <= 0 || T <= 0 || neq < 1 || neq > FVM_MAX_NEQ || (Eb > 0 && !bedges) || !stats)
        return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_plan_selftest: bad arguments");
    if (tile_triangles < 64 || tile_triangles > 4096 || tile_triangles % 64) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_plan_selftest: bad tile size");
    fvm_ctx ctx;
    fvm_ctx* h = &ctx;
    h->N = N;
    h->T = T;
    h->Eb = Eb;
    h->neq = neq;
    h->h_xy.assign(xy, xy + 2 * N);
    h->h_tri.resize(3 * T);
    for (int64_t i = 0; i < 3 * T; ++i) {
        h->h_tri[i] = tri[i] - index_base;
        if (h->h_tri[i] < 0 || h->h_tri[i] >= N) return fvm_fail(nullptr, FVM_ERR_ARG, "fvm_plan_selftest: triangle vertex out of range");
    }
    h->h_bedge.resize(2 * Eb);
    for (int64_t i = 0; i < 2 * Eb; ++i) h->h_bedge[i] = bedges[i] - index_base;
    for (int v = 0; v < neq; ++v) {
        h->h_nkind[v].assign(N, 0);
        if (node_kind) h->h_nkind[v].assign(node_kind + (size_t)v * N, node_kind + (size_t)(v + 1) * N);
        h->h_nfidx[v].assign(N, 0);
        h->h_ekind[v].assign(Eb, 0);
        h->h_efidx[v].assign(Eb, 0);
    }
    const int TT = tile_triangles;
    HostPlan P;
    int32_t rc = plan_host(h, TT, P);
    if (rc) return fvm_fail(nullptr, rc, h->err);
    auto bad = [&](const std::string& what) { return fvm_fail(nullptr, FVM_ERR_STATE, "fvm_plan_selftest: " + what); };
    // permutations
    {
        std::vector<uint8_t> seen(T, 0);
        for (int64_t t = 0; t < T; ++t) {
            const int32_t o = h->tri_old_of_new[t];
            if (o < 0 || o >= T || seen[o]) return bad("triangle order is not a permutation");
            seen[o] = 1;
        }
        std::vector<uint8_t> seen_n(N, 0);
        for (int64_t g = 0; g < N; ++g) {
            const int32_t o = h->node_old_of_new[g];
            if (o < 0 || o >= N || seen_n[o] || h->node_new_of_old[o] != g) return bad("node renumbering is not a permutation");
            seen_n[o] = 1;
        }
    }
    // tile ranges tile the vertex range
    int64_t cursor = 0, gather_entries = 0;
    std::vector<int32_t> touch(N, 0);  // number of tiles that hold a native node as a local node
    for (int64_t b = 0; b < P.n_tiles; ++b) {
        const int32_t node0 = P.tile_node0[b], nint = P.tile_nint[b], nown = P.tile_nown[b], nloc = P.tile_nloc[b];
        if (node0 != cursor || nint < 0 || nint > nown || nown > nloc || nloc > P.max_nloc || nloc >= 65535) return bad("tile node ranges are inconsistent");
        cursor += nown;
        if (P.tile_ext0[b + 1] - P.tile_ext0[b] != nloc - nown) return bad("external node count mismatch");
        const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
        auto native_of = [&](int32_t l) { return l < nown ? node0 + l : P.ext_ids[P.tile_ext0[b] + (l - nown)]; };
        for (int32_t l = 0; l < nloc; ++l) {
            const int32_t g = native_of(l);
            if (g < 0 || g >= P.n_vertices) return bad("local node maps outside the vertex range");
            if (l >= nown && g >= node0 && g < node0 + nown) return bad("an own node is listed as external");
            touch[g]++;
        }
        // triangles: local ids decode to the triangle's native vertices
        for (int64_t nt = t0; nt < t1; ++nt) {
            const ushort4 q = P.tri_loc[nt];
            const int32_t l3[3] = {q.x, q.y, q.z};
            for (int r = 0; r < 3; ++r) {
                if (l3[r] >= nloc) return bad("tile-local vertex id out of range");
                const int32_t want = h->node_new_of_old[h->h_tri[3 * (int64_t)h->tri_old_of_new[nt] + r]];
                if (native_of(l3[r]) != want || P.tri_native[3 * nt + r] != want) return bad("tile-local vertex id decodes to the wrong node");
            }
        }
        // gather list: every (triangle, slot) exactly once, under the node it belongs to, ascending triangle order
        const uint16_t* iptr = P.inc_ptr.data() + P.tile_loc0[b];
        const uint16_t* inc = P.inc.data() + (size_t)3 * TT * b;
        if (iptr[0] != 0 || iptr[nloc] != 3 * (t1 - t0)) return bad("gather list does not cover the tile");
        std::vector<uint8_t> hit(3 * (t1 - t0), 0);
        for (int32_t l = 0; l < nloc; ++l) {
            int32_t prev = -1;
            for (int32_t e = iptr[l]; e < iptr[l + 1]; ++e) {
                const int32_t lt = inc[e] >> 2, slot = inc[e] & 3;
                if (lt >= t1 - t0 || slot > 2 || lt <= prev) return bad("gather list entry out of range or out of order");
                prev = lt;
                const ushort4 q = P.tri_loc[t0 + lt];
                if ((slot == 0 ? q.x : slot == 1 ? q.y : q.z) != l) return bad("gather list entry points at another node");
                if (hit[3 * lt + slot]++) return bad("gather list entry repeated");
                ++gather_entries;
            }
        }
    }
    if (cursor != P.n_vertices) return bad("tile ranges do not cover the vertices");
    // interior nodes belong to one tile; interface nodes own one partial slot per touching tile and live edge end
    std::vector<int32_t> edge_ends(N, 0);
    for (const BndEdge& e : P.bnd) {
        edge_ends[e.v[e.pi]]++;
        edge_ends[e.v[e.pj]]++;
    }
    std::vector<uint8_t> slot_used(P.n_partial, 0);
    int32_t ifc = 0;
    for (int64_t b = 0; b < P.n_tiles; ++b) {
        for (int32_t l = 0; l < P.tile_nown[b]; ++l) {
            const int32_t g = P.tile_node0[b] + l;
            if (l < P.tile_nint[b]) {
                if (touch[g] != 1 || edge_ends[g] != 0) return bad("an interior node is shared or carries a live boundary edge");
            } else {
                if (ifc >= P.n_ifc || P.ifc_node[ifc] != g) return bad("interface list is not in tile order");
                if (P.ifc_pptr[ifc + 1] - P.ifc_pptr[ifc] != touch[g] + edge_ends[g]) return bad("partial slots of an interface node do not match its tiles and edges");
                ++ifc;
            }
        }
        // the tile's partial positions: own interface nodes, then external nodes
        const int32_t n_if_local = (P.tile_nown[b] - P.tile_nint[b]) + (P.tile_nloc[b] - P.tile_nown[b]);
        if (P.tile_pp0[b + 1] - P.tile_pp0[b] != n_if_local) return bad("partial position list has the wrong length");
        for (int32_t q = 0; q < n_if_local; ++q) {
            const int32_t l = P.tile_nint[b] + q;
            const int32_t g = l < P.tile_nown[b] ? P.tile_node0[b] + l : P.ext_ids[P.tile_ext0[b] + (l - P.tile_nown[b])];
            const int32_t k = (int32_t)(std::lower_bound(P.ifc_node.begin(), P.ifc_node.end(), g) - P.ifc_node.begin());
            if (k >= P.n_ifc || P.ifc_node[k] != g) return bad("a shared node is not an interface node");
            const int32_t pos = P.ppos[P.tile_pp0[b] + q];
            if (pos < P.ifc_pptr[k] || pos >= P.ifc_pptr[k + 1] || slot_used[pos]++) return bad("partial slot outside its node's range or used twice");
        }
    }
    if (ifc != P.n_ifc) return bad("interface node count mismatch");
    for (const BndEdge& e : P.bnd)
        for (int q = 0; q < 2; ++q) {
            const int32_t g = e.v[q == 0 ? e.pi : e.pj], pos = q == 0 ? e.slot_i : e.slot_j;
            const int32_t k = (int32_t)(std::lower_bound(P.ifc_node.begin(), P.ifc_node.end(), g) - P.ifc_node.begin());
            if (k >= P.n_ifc || P.ifc_node[k] != g) return bad("a live boundary edge ends at a node that is not an interface node");
            if (pos < P.ifc_pptr[k] || pos >= P.ifc_pptr[k + 1] || slot_used[pos]++) return bad("boundary-edge partial slot outside its node's range or used twice");
        }
    for (int64_t s = 0; s < P.n_partial; ++s)
        if (!slot_used[s]) return bad("an allocated partial slot has no writer");
    // tile packs of the streaming kernel: every contribution is gathered exactly once by the node it belongs to, and
    // neither the write groups (16 consecutive triangles, one slot) nor the read groups (16 consecutive local nodes,
    // one gather row) touch an 8-byte bank pair twice
    if (TT <= FVM_STREAM_MAX_TT) {
        std::vector<double> xyn(2 * N);
        std::vector<uint8_t> kn((size_t)neq * N, 0);
        for (int64_t g = 0; g < N; ++g) {
            xyn[2 * g] = xy[2 * (int64_t)h->node_old_of_new[g]];
            xyn[2 * g + 1] = xy[2 * (int64_t)h->node_old_of_new[g] + 1];
        }
        fvm_rawvec<uint8_t> packs;  // every byte is written by build_tile_packs (memset + fields, tile by tile)
        std::vector<int4> pdir;
        int32_t cap = 0;
        PhaseTimer tmp;
        rc = build_tile_packs(h, P, TT, xyn.data(), kn.data(), nullptr, packs, pdir, cap);
        tmp.lap("tile packs (selftest)");
        if (rc) return fvm_fail(nullptr, rc, h->err);
        for (int64_t b = 0; b < P.n_tiles; ++b) {
            const uint8_t* base = packs.data() + ((size_t)(uint32_t)pdir[2 * b].x << 4);
            TilePackHdr H;
            std::memcpy(&H, base, sizeof(H));
            const int64_t t0 = b * TT, t1 = std::min<int64_t>(T, t0 + TT);
            if (H.node0 != P.tile_node0[b] || H.nint != P.tile_nint[b] || H.nown != P.tile_nown[b] || H.nloc != P.tile_nloc[b] ||
                H.ntri != t1 - t0 || H.bytes != pdir[2 * b].y || H.bytes > cap || (H.bytes & 15) || pdir[2 * b].z != H.node0 ||
                pdir[2 * b].w != H.nown || pdir[2 * b + 1].y != H.nloc - H.nown)
                return bad("tile pack header does not match the plan");
            const ushort4* tq = reinterpret_cast<const ushort4*>(base + H.off_tri);
            const double* pxy = reinterpret_cast<const double*>(base + H.off_xy);
            const uint16_t* urow = reinterpret_cast<const uint16_t*>(base + H.off_urow);
            const uint16_t* list = reinterpret_cast<const uint16_t*>(base + H.off_list);
            for (int32_t l = 0; l < H.nloc; ++l) {
                const int32_t g = l < H.nown ? H.node0 + l : P.ext_ids[P.tile_ext0[b] + (l - H.nown)];
                if (pxy[2 * l] != xyn[2 * (size_t)g] || pxy[2 * l + 1] != xyn[2 * (size_t)g + 1]) return bad("tile pack coordinates are wrong");
            }
            // write groups
            std::vector<int32_t> owner((size_t)3 * TT, -1);  // contribution slot -> 3 * lt + slot
            for (int32_t lt = 0; lt < H.ntri; ++lt) {
                if (tq[lt].x != P.tri_loc[t0 + lt].x || tq[lt].y != P.tri_loc[t0 + lt].y || tq[lt].z != P.tri_loc[t0 + lt].z) return bad("tile pack triangle ids are wrong");
                for (int sl = 0; sl < 3; ++sl) {
                    const int32_t pos = sl * TT + (lt & ~15) + ((tq[lt].w >> (4 * sl)) & 15);
                    if (owner[pos] >= 0) return bad("two contributions of a write group share a slot (bank conflict)");
                    owner[pos] = 3 * lt + sl;
                }
            }
            // read groups: one gather row of half a 32-node unit
            std::vector<uint8_t> got((size_t)3 * TT, 0);
            if (H.nunit != (H.nloc + 31) / 32 || urow[0] != 0) return bad("tile pack unit table is wrong");
            for (int32_t q = 0; q < H.nunit; ++q)
                for (int32_t r = urow[q]; r < urow[q + 1]; ++r)
                    for (int half = 0; half < 2; ++half) {
                        int used[16] = {0}, padcode = -1;
                        for (int i = 16 * half; i < 16 * half + 16; ++i) {
                            const int32_t code = list[(size_t)r * 32 + i];
                            if (code & 7) return bad("gather code is not a multiple of 8");
                            const int32_t pos = code >> 3, l = 32 * q + i;
                            if (pos >= 3 * TT) {  // padded entry: one of the 16 zero words
                                if (pos >= 3 * TT + 16 || (padcode >= 0 && padcode != code)) return bad("bad padded gather entry");
                                padcode = code;
                                continue;
                            }
                            if (l >= H.nloc || owner[pos] < 0 || got[pos]++) return bad("gather entry points at no contribution or at one already gathered");
                            const int32_t lt = owner[pos] / 3, slot = owner[pos] % 3;
                            if ((slot == 0 ? tq[lt].x : slot == 1 ? tq[lt].y : tq[lt].z) != l) return bad("gather entry belongs to another node");
                            if (used[pos & 15]++) return bad("two entries of a read group share a bank pair");
                        }
                        if (padcode >= 0 && used[(padcode >> 3) & 15]) return bad("padded entries share a bank pair with a real entry");
                    }
            for (int32_t lt = 0; lt < H.ntri; ++lt)
                for (int sl = 0; sl < 3; ++sl)
                    if (!got[sl * TT + (lt & ~15) + ((tq[lt].w >> (4 * sl)) & 15)]) return bad("a contribution is never gathered");
        }
    }
    stats[0] = P.n_tiles;
    stats[1] = P.n_vertices;
    stats[2] = P.n_ifc;
    stats[3] = P.n_partial;
    stats[4] = (int64_t)P.ext_ids.size();
    stats[5] = P.max_nloc;
    stats[6] = (int64_t)P.bnd.size();
    stats[7] = gather_entries;
    return FVM_OK;
}

extern "C" int32_t fvm_finalize(fvm_handle h, int32_t tile_triangles, int32_t geometry_mode) {
    NOT_FINAL(h);
    FVM_CUDA(h, cudaSetDevice(h->device));
    const int64_t N = h->N, T = h->T, Eb = h->Eb;
    const int neq = h->neq;
    // default tile size (measured on B200, 4096^2): kernels that stream the full 21-component SoA are
    // register-limited to 3 CTAs/SM and like big tiles; the reduced stream wants 8 small CTAs per SM
    const bool full_flux = h->flux.model == FVM_FLUX_DIFF_POWER || h->flux.model == FVM_FLUX_ADVDIFF || h->flux.model == FVM_FLUX_KELLER_SEGEL;
    // (systems: 768 measured best for the 2-species Keller-Segel kernel: 1.20 ms vs 1.45 ms at 1024)
    // recompute mode (streaming kernel, fvm_rhs_stream.cu): 768 = three full triangle iterations of its 256 consumer threads
    // and 14 of 16 node-pass slots, 3 CTAs/SM with one contribution buffer (see launch_stream_threads); systems: 512
    int TT = tile_triangles > 0 ? tile_triangles
                                : (geometry_mode == 1 ? (neq >= 2 ? 512 : 768) : (neq >= 2 ? 768 : (!full_flux ? 512 : 1024)));
    if (tile_triangles <= 0)
        if (const char* e = getenv("FVM_TILE_TRIANGLES")) TT = atoi(e);
    FVM_REQUIRE(h, TT >= 64 && TT <= 4096 && TT % 64 == 0, "fvm_finalize: tile_triangles must be a multiple of 64 in 64..4096");
    FVM_REQUIRE(h, geometry_mode == 0 || geometry_mode == 1, "fvm_finalize: geometry_mode must be 0 or 1");
    if (h->flux.model == FVM_FLUX_DIFF_TABLE && h->h_dtab.empty())
        return fvm_fail(h, FVM_ERR_STATE, "fvm_finalize: table flux model without fvm_set_flux_table");
    h->geometry_mode = geometry_mode;
    if (const char* e = getenv("FVM_STREAM_THREADS")) {
        const int v = atoi(e);
        if (v == 192 || v == 224 || v == 256 || v == 384 || v == 512) h->stream_threads = v;
    }
    if (const char* e = getenv("FVM_STREAM_OCC")) h->stream_occ = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 3);  // experiment knob
    const double* xy = h->h_xy.data();
    const int32_t* tri = h->h_tri.data();

    HostPlan P;
    PhaseTimer tm;
    int32_t rc_plan = plan_host(h, TT, P);
    if (rc_plan) return rc_plan;
    tm.lap("plan_host total");
    const int64_t n_tiles = P.n_tiles, tpad = P.tpad, n_partial = P.n_partial;
    const int32_t n_vertices = P.n_vertices, n_ifc = P.n_ifc, max_nloc = P.max_nloc;
    const int32_t* told = h->tri_old_of_new.data();
    std::vector<int32_t>&tile_node0 = P.tile_node0, &tile_nint = P.tile_nint, &tile_nown = P.tile_nown, &tile_nloc = P.tile_nloc;
    std::vector<int32_t>&tile_ext0 = P.tile_ext0, &tile_loc0 = P.tile_loc0, &tile_pp0 = P.tile_pp0;
    fvm_rawvec<ushort4>& tri_loc = P.tri_loc;
    fvm_rawvec<int32_t>& tri_native = P.tri_native;
    std::vector<int32_t>&ext_ids = P.ext_ids, &ifc_node = P.ifc_node, &ifc_pptr = P.ifc_pptr, &ppos = P.ppos;
    std::vector<uint16_t>& inc_ptr = P.inc_ptr;
    fvm_rawvec<uint16_t>& inc = P.inc;
    std::vector<BndEdge>& bnd = P.bnd;
    std::vector<double>& dbnd_live = P.dbnd_live;
    // ---- 6. upload --------------------------------------------------------------------------
    DevMesh& m = h->dm;
    m.neq = neq;
    m.n_nodes = (int32_t)N;
    m.n_tris = (int32_t)T;
    m.n_tiles = (int32_t)n_tiles;
    m.tile_tris = TT;
    m.tpad = tpad;
    {   // distance (in tiles) between a running CTA and the CTAs about to be scheduled: SMs x resident CTAs
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const char* e = getenv("FVM_PF_AHEAD");
        m.pf_ahead = e ? atoi(e) : 2 * sms;  // measured at 4096^2: 0.562 ms without, 0.533 at 148..296, 0.550 at 1184
    }
    int32_t rc;
#define UP(dst, vec)                                               \
    do {                                                           \
        std::remove_const<std::remove_pointer<decltype(dst)>::type>::type* p__ = nullptr; \
        if ((rc = fvm_dev_upload(h, &p__, vec))) return rc;        \
        dst = p__;                                                 \
    } while (0)
    UP(m.tri_loc, tri_loc);
    {
        std::vector<int4> meta(2 * n_tiles);
        for (int64_t b = 0; b < n_tiles; ++b) {
            meta[2 * b] = make_int4(tile_node0[b], tile_nint[b], tile_nown[b], tile_nloc[b]);
            meta[2 * b + 1] = make_int4(tile_ext0[b], tile_loc0[b], tile_pp0[b], (int)std::min<int64_t>(TT, T - b * TT));
        }
        UP(m.tile_meta, meta);
    }
    h->h_tile_node0 = tile_node0;
    h->h_tile_nint = tile_nint;
    h->h_ifc_node = ifc_node;
    h->h_bnd = bnd;
    h->h_tile_nown = tile_nown;
    h->h_tile_ext0 = tile_ext0;
    h->h_ext_ids = ext_ids;
    UP(m.tile_node0, tile_node0);
    UP(m.tile_nint, tile_nint);
    UP(m.tile_nown, tile_nown);
    UP(m.tile_nloc, tile_nloc);
    UP(m.tile_ext0, tile_ext0);
    UP(m.tile_loc0, tile_loc0);
    UP(m.tile_pp0, tile_pp0);
    UP(m.ext_ids, ext_ids);
    UP(m.inc_ptr, inc_ptr);
    UP(m.inc, inc);
    UP(m.ppos, ppos);
    UP(m.ifc_node, ifc_node);
    UP(m.ifc_pptr, ifc_pptr);
    m.n_ifc = n_ifc;
    m.n_partial = n_partial;
    if ((rc = fvm_dev_alloc(h, &m.partial, (size_t)n_partial * neq))) return rc;
    // per-node arrays in native order
    fvm_rawvec<double> xyn(2 * N);  // (filled completely by the parallel loop below)
    fvm_rawvec<uint8_t> kn((size_t)neq * N);
    {
        fvm_rawvec<int32_t> fn((size_t)neq * N);
        std::vector<int32_t> dir;
#pragma omp parallel for schedule(static)
        for (int64_t g = 0; g < N; ++g) {
            const int32_t o = h->node_old_of_new[g];
            xyn[2 * g] = xy[2 * o];
            xyn[2 * g + 1] = xy[2 * o + 1];
            for (int v = 0; v < neq; ++v) {
                kn[(size_t)v * N + g] = g < n_vertices ? h->h_nkind[v][o] : (uint8_t)3;  // 3: not a vertex
                fn[(size_t)v * N + g] = h->h_nfidx[v][o];
            }
        }
        for (int v = 0; v < neq; ++v)
            for (int64_t g = 0; g < n_vertices; ++g)
                if (kn[(size_t)v * N + g] == FVM_NODE_DIRICHLET) {
                    dir.push_back((int32_t)g);
                    dir.push_back(v);
                }
        UP(m.xy, xyn);
        UP(m.kind, kn);
        UP(m.fidx, fn);
        h->n_dir = (int32_t)(dir.size() / 2);
        if ((rc = fvm_dev_upload(h, &h->d_dir_nodes, dir))) return rc;
        if (!h->h_srctab.empty()) {
            std::vector<double> sn((size_t)N * neq);
#pragma omp parallel for schedule(static)
            for (int64_t g = 0; g < N; ++g)
                for (int v = 0; v < neq; ++v) sn[g * neq + v] = h->h_srctab[(int64_t)h->node_old_of_new[g] * neq + v];
            UP(m.src_tab, sn);
        }
    }
    std::vector<double> dt;
    if (!h->h_dtab.empty()) {
        dt.assign((size_t)3 * tpad, 0.0);
#pragma omp parallel for schedule(static)
        for (int64_t nt = 0; nt < T; ++nt)
            for (int e = 0; e < 3; ++e) dt[(size_t)e * tpad + nt] = h->h_dtab[3 * (int64_t)told[nt] + e];
        UP(m.dtab, dt);
    }
    tm.lap("upload + node arrays");
    // tile packs of the streaming recompute kernel (geometry_mode 1; big template tiles keep the plain tile kernel)
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (geometry_mode == 1 && TT <= FVM_STREAM_MAX_TT && !getenv("FVM_NO_STREAM")) {
        fvm_rawvec<uint8_t> packs;  // every byte is written by build_tile_packs (memset + fields, tile by tile)
        std::vector<int4> pdir;
        if ((rc = build_tile_packs(h, P, TT, xyn.data(), kn.data(), dt.empty() ? nullptr : dt.data(), packs, pdir, h->pack_cap))) return rc;
        if ((rc = fvm_dev_upload(h, &h->d_packs, packs))) return rc;
        if ((rc = fvm_dev_upload(h, &h->d_pack_dir, pdir))) return rc;
        h->pack_bytes_total = (int64_t)packs.size();
        FVM_CUDA(h, cudaStreamSynchronize(h->stream));  // the host vectors die at the end of this block
    }
    std::vector<double>().swap(dt);
    tm.lap("tile packs");
    UP(m.cond, h->h_cond);
    if ((rc = fvm_dev_upload(h, &h->d_node_old_of_new, h->node_old_of_new))) return rc;
    if ((rc = fvm_dev_upload(h, &h->d_node_new_of_old, h->node_new_of_old))) return rc;
    h->n_bnd_live = (int32_t)bnd.size();
    if ((rc = fvm_dev_upload(h, &h->d_bnd, bnd))) return rc;
    if (!dbnd_live.empty() && (rc = fvm_dev_upload(h, &h->d_dbnd, dbnd_live))) return rc;
    {
        double* vol = nullptr;
        if ((rc = fvm_dev_alloc(h, &vol, (size_t)N))) return rc;
        m.vol = vol;
    }
    h->max_nloc = max_nloc;

    // ---- 7. geometry on the device ----------------------------------------------------------
    int32_t* d_tri_native = nullptr;
    FVM_CUDA(h, cudaMalloc((void**)&d_tri_native, sizeof(int32_t) * 3 * T));
    FVM_CUDA(h, cudaMemcpyAsync(d_tri_native, tri_native.data(), sizeof(int32_t) * 3 * T, cudaMemcpyHostToDevice, h->stream));
    h->d_tri_native = d_tri_native;  // kept: assembly and geometry export read it
    h->allocs.push_back(d_tri_native);
    rc = fvm_launch_geometry(h, d_tri_native);
    if (rc) return rc;
    rc = fvm_launch_volumes(h);
    if (rc) return rc;
    if (h->d_packs) {
        if ((rc = fvm_stream_fill_vinv(h))) return rc;
        h->packs_ready = true;
    }
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    tm.lap("device geometry + volumes");

    h->stats[0] = n_tiles;
    h->stats[1] = TT;
    h->stats[2] = n_vertices;
    h->stats[3] = n_ifc;
    h->stats[4] = n_partial;
    h->stats[5] = (int64_t)ext_ids.size();
    h->stats[6] = max_nloc;
    h->stats[7] = h->n_bnd_live;
    h->stats[8] = h->n_dir;
    h->stats[9] = h->smem_rhs;
    h->finalized = true;
    // the big host copies are no longer needed
    std::vector<double>().swap(h->h_dtab);
    std::vector<double>().swap(h->h_srctab);
    return FVM_OK;
}

extern "C" int32_t fvm_get_permutation(fvm_handle h, int32_t* node_perm, int32_t* tri_perm) {
    if (!h) return FVM_ERR_ARG;
    if (!h->finalized) return fvm_fail(h, FVM_ERR_STATE, "fvm_get_permutation: finalize first");
    if (node_perm) std::memcpy(node_perm, h->node_old_of_new.data(), sizeof(int32_t) * h->N);
    if (tri_perm) std::memcpy(tri_perm, h->tri_old_of_new.data(), sizeof(int32_t) * h->T);
    return FVM_OK;
}

extern "C" int32_t fvm_get_stats(fvm_handle h, int64_t* stats) {
    if (!h || !stats) return FVM_ERR_ARG;
    for (int i = 0; i < 16; ++i) stats[i] = h->stats[i];
    stats[9] = h->smem_rhs;
    stats[15] = h->pipe_choice[0] + 4 * h->pipe_choice[1];  // host-buffer schedule chosen for fvm_rhs / fvm_spmv: 0 undecided, 1 pipeline, 2 plain
    return FVM_OK;
}
