// Internal context of libfvmcuda.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/fvmcuda.h"

#define FVM_MAX_NEQ 4
#define FVM_MAX_PARAMS 16
#define FVM_MAX_COND_FN 64
#define FVM_NGEO 21  // s1..s9, 3 cv-edge midpoints (x,y), 3 length-scaled normals (x,y)

struct CondFn {
    int32_t id;
    double p[4];
};

// Device-side view of everything the kernels need; passed by value (kernel parameter space).
struct DevMesh {
    int32_t neq;
    int32_t n_nodes;       // all points
    int32_t n_vertices;    // points that are triangle vertices; native ids [0, n_vertices)
    int32_t n_tris;
    int32_t n_tiles;
    int32_t tile_tris;     // TT
    int32_t pf_ahead;      // recompute kernels prefetch the lists of tile + pf_ahead into L2 (0 = off)
    int64_t tpad;          // n_tiles * TT : stride between geometry components
    // per triangle (native order, padded to tpad)
    const ushort4* tri_loc;  // 3 tile-local vertex ids (+ spare)
    const double* geo;       // [FVM_NGEO][tpad]
    const double* dtab;      // [3][tpad] tabulated D at cv-edge midpoints (or null)
    // per tile
    const int4* tile_meta;       // 2 x int4 per tile: {node0, nint, nown, nloc}, {ext0, loc0, pp0, ntri}
    const int32_t* tile_node0;   // first own node (native id)
    const int32_t* tile_nint;    // interior nodes
    const int32_t* tile_nown;    // interior + owned interface nodes
    const int32_t* tile_nloc;    // own + external interface nodes
    const int32_t* tile_ext0;    // offset into ext_ids
    const int32_t* tile_loc0;    // offset into inc_ptr (per local node, +1 sentinel per tile)
    const int32_t* tile_pp0;     // offset into ppos (per interface local)
    const int32_t* ext_ids;      // native ids of external interface nodes
    const uint16_t* inc_ptr;     // per local node offset into the tile's incidence list
    const uint16_t* inc;         // [n_tiles][3*TT] entries (local_tri << 2 | slot)
    const int32_t* ppos;         // partial-buffer slot of each interface local
    // per node (native order)
    const double* xy;            // interleaved
    const double* vol;           // control-volume areas
    const uint8_t* kind;         // [neq][n_nodes]
    const int32_t* fidx;         // [neq][n_nodes]
    const double* src_tab;       // [n_nodes][neq] or null
    // interface nodes
    int32_t n_ifc;
    const int32_t* ifc_node;     // native id
    const int32_t* ifc_pptr;     // [n_ifc+1] offsets into partial (units of neq doubles)
    double* partial;             // [n_partial][neq]
    int64_t n_partial;
    // condition function table
    const CondFn* cond;          // [neq][FVM_MAX_COND_FN]
};

// Tile pack of the streaming (recompute-geometry) RHS kernel (fvm_rhs_stream.cu): everything static a tile
// needs -- local vertex ids, vertex coordinates (own + external), 1/V, condition kinds, partial slots and the
// node gather list -- as ONE 16-byte aligned record, fetched by a single bulk async copy (cp.async.bulk).
struct TilePackHdr {  // 64 bytes, section offsets in bytes from the start of the pack
    int32_t node0, nint, nown, nloc;
    int32_t ntri, flags, nunit, bytes;  // flags bit 0: every interior node is free in every species
    int32_t off_tri, off_xy, off_vinv, off_kind;
    int32_t off_ppos, off_urow, off_list, off_dtab;
};
#define FVM_STREAM_MAX_TT 2688  // gather codes are 16-bit byte offsets into the contribution planes: 3*TT*8 + 8 <= 65536

struct BndEdge {  // one live boundary edge (native ids); tiny count, AoS is fine
    int32_t v[3];      // stored vertex triple of the adjacent triangle
    int32_t pi, pj;    // positions of i and j inside v
    int32_t slot_i, slot_j;  // partial-buffer slots
    int32_t orig;      // caller's edge index
    double px, py, qx, qy;
    uint8_t kind[FVM_MAX_NEQ];
    int32_t fidx[FVM_MAX_NEQ];
};

struct FluxParams {
    int32_t model;
    int32_t nparams;
    double p[FVM_MAX_PARAMS];
};

struct SourceParams {
    int32_t model;
    int32_t nparams;
    double p[FVM_MAX_PARAMS];
};

struct Csr {
    int64_t nnz = 0;
    int32_t n = 0;
    int32_t max_row = 0;
    bool pattern = false;
    int32_t* n2t_ptr = nullptr;  // node -> incident (triangle << 2 | slot), native ids
    int32_t* n2t = nullptr;
    double* diag_inv = nullptr;  // Jacobi preconditioner of the (scaled) system
    int32_t chunk_rows = 0;      // SpMV: rows per CTA
    int32_t chunk_smem = 0;
    // tile-local SpMV: interior rows of a tile are contiguous in native order, so their val/col span is
    // contiguous too; col16 holds the tile-local column of every entry of an interior row
    uint16_t* col16 = nullptr;
    int32_t* tail_rows = nullptr;  // interface rows + points that are not vertices (generic CSR path)
    int32_t n_tail = 0;
    int32_t tile_prod_cap = 0, tile_max_nint = 0, tile_smem = 0;
    int32_t use_tile_spmv = 1;
    // sliced-ELL copy of the interior rows (32 rows per slice, entry k of the 32 rows contiguous)
    int32_t* tile_slice0 = nullptr;  // [n_tiles + 1] first slice of each tile
    int32_t* sell_ptr = nullptr;     // [n_slices + 1] offset of each slice (units of entries)
    double* sell_val = nullptr;
    uint16_t* sell_col = nullptr;
    int64_t sell_entries = 0;
    int32_t n_slices = 0;
    // sliced-ELL copy of the tail rows (global int32 columns)
    int32_t* tsell_ptr = nullptr;
    double* tsell_val = nullptr;
    int32_t* tsell_col = nullptr;
    int32_t n_tslices = 0;
    int32_t* rowptr = nullptr;  // native numbering
    int32_t* col = nullptr;
    double* val = nullptr;
    double* b = nullptr;
    double* rowscale = nullptr;  // -V (free rows) / 1 (identity rows), for the symmetrised PCG
    bool assembled = false;
    int32_t template_id = -1;
};

struct fvm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool finalized = false;
    bool time_dependent = false;  // some registered condition function reads t
    int32_t neq = 1;
    int32_t h_index_base = 0;
    void* graph_exec = nullptr;
    int32_t* d_tri_native = nullptr;  // [T][3] native node ids per native triangle
    int64_t N = 0, T = 0, Eb = 0;
    int32_t geometry_mode = 0;
    // host copies of the caller's mesh (caller order, 0-based)
    std::vector<double> h_xy;
    std::vector<int32_t> h_tri;
    std::vector<int32_t> h_bedge;
    std::vector<uint8_t> h_ekind[FVM_MAX_NEQ];
    std::vector<int32_t> h_efidx[FVM_MAX_NEQ];
    std::vector<uint8_t> h_nkind[FVM_MAX_NEQ];
    std::vector<int32_t> h_nfidx[FVM_MAX_NEQ];
    std::vector<double> h_dtab, h_dbnd, h_srctab;
    std::vector<int32_t> h_edge_tri, h_edge_rot;  // adjacent triangle (caller id) / rotation per boundary edge
    std::vector<CondFn> h_cond;  // [neq][FVM_MAX_COND_FN]
    FluxParams flux{};
    SourceParams source{};
    // permutations
    std::vector<int32_t> node_new_of_old, node_old_of_new, tri_old_of_new, tri_new_of_old;
    int32_t* d_node_old_of_new = nullptr;
    int32_t* d_node_new_of_old = nullptr;
    // device mesh
    DevMesh dm{};
    std::vector<void*> allocs;  // every cudaMalloc, freed in destroy
    BndEdge* d_bnd = nullptr;
    int32_t n_bnd_live = 0;
    double* d_dbnd = nullptr;  // [n_bnd_live][2] tabulated D at the quarter points
    int32_t n_dir = 0;         // Dirichlet (node,species) pairs for the callback kernel
    int32_t* d_dir_nodes = nullptr;
    int32_t smem_rhs = 0;
    std::map<const void*, int32_t> smem_configured;  // kernel -> opted-in dynamic smem bytes
    int32_t max_nloc = 0;
    // scratch state vectors (native order), lazily sized
    double *d_u = nullptr, *d_du = nullptr, *d_io = nullptr;
    // linear path
    Csr csr;
    double* d_work[12] = {nullptr};
    double* d_red = nullptr;  // reduction scratch
    double* d_dotpart = nullptr;  // per-CTA partial dot products of the fused PCG operator
    int32_t dotpart_n = 0;
    int64_t stats[16] = {0};
    // optional per-kernel timing of the dominant kernels (bench.py's roofline leg)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
    int64_t prof_used = 0;
    // sparse Jacobian (fvm_jacobian.cu): block values on the CSR pattern, boundary edges by node
    double* jac_val = nullptr;
    void* jac_edges = nullptr;
    int32_t *jac_bn_of = nullptr, *jac_bn_ptr = nullptr, *jac_bn_items = nullptr;
    bool jac_ready = false;
    // sharding (fvm_shard.cu)
    void* shard = nullptr;
    bool halo_ready = false;
    bool overlap = false;            // exchange overlapped with the tiles that touch no ghost node
    cudaStream_t comm_stream = nullptr;   // communication stream (owned by the shard state)
    cudaStream_t launch_stream = nullptr; // stream the kernel launchers use (compute stream by default)
    int32_t* d_tile_order = nullptr; // independent tiles first, then the halo-dependent ones
    int32_t n_tiles_indep = 0;
    std::vector<int32_t> h_tile_node0, h_tile_nown, h_tile_ext0, h_ext_ids;  // host copies for the classification
    std::vector<uint8_t> h_ghost;  // caller order: 1 = ghost node owned by another rank
    int32_t rank = 0, nranks = 1;
    // host-buffer pipeline of fvm_rhs (fvm_pipe.cu): caller-order bands copied in, tiles launched as soon as
    // their nodes have arrived, finished bands copied out while later bands are still being copied in
    std::vector<int32_t> h_tile_nint, h_ifc_node;
    std::vector<BndEdge> h_bnd;       // host copy of the live boundary-edge records
    void* pipe = nullptr;       // plan of fvm_rhs
    void* pipe_spmv = nullptr;  // plan of fvm_spmv
    const int32_t* pipe_list = nullptr;  // explicit tile list of fvm_launch_rhs_part(part = 4)
    int32_t pipe_off = 0, pipe_count = 0;
    int32_t pipe_choice[2] = {0, 0};  // schedule chosen by the one-off timing (fvm_rhs, fvm_spmv): 0 undecided, 1 pipeline, 2 plain
    // streaming recompute kernel (fvm_rhs_stream.cu): per-tile packs + directory
    uint8_t* d_packs = nullptr;
    int4* d_pack_dir = nullptr;   // 2 x int4 per tile: {pack offset / 16, pack bytes, node0, nown}, {ext0, next, 0, 0}
    int32_t pack_cap = 0;         // largest pack (bytes)
    int64_t pack_bytes_total = 0;
    bool packs_ready = false;
    int32_t stream_threads = 0;    // consumer threads per CTA of the streaming kernel (0: chosen per flux model, FVM_STREAM_THREADS)
    int32_t stream_occ = 0;        // resident CTAs per SM its register budget is planned for (256-thread variant); 0: 3, systems 2
    std::map<const void*, int32_t> occ_cache;  // kernel -> resident CTAs per SM at the configured shared memory
    int32_t sm_count = 148;
};

// ---- helpers -----------------------------------------------------------------
int32_t fvm_fail(fvm_ctx* h, int32_t code, const std::string& msg);
#define FVM_CUDA(h, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            cudaGetLastError(); /* a reported error must not fail the next, unrelated call */     \
            return fvm_fail((h), FVM_ERR_CUDA,                                                    \
                            std::string(#call) + ": " + cudaGetErrorString(e__));                \
        }                                                                                         \
    } while (0)
#define FVM_REQUIRE(h, cond, msg)                                 \
    do {                                                          \
        if (!(cond)) return fvm_fail((h), FVM_ERR_ARG, (msg));    \
    } while (0)

template <class Tp>
int32_t fvm_dev_alloc(fvm_ctx* h, Tp** p, size_t count) {
    void* q = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(Tp);
    FVM_CUDA(h, cudaMalloc(&q, bytes));
    h->allocs.push_back(q);
    *p = (Tp*)q;
    return FVM_OK;
}
// std::vector whose resize() leaves trivially constructible elements uninitialised: the big host staging buffers of
// fvm_finalize (tile packs: 1 GB at 4096^2) are first touched by the OpenMP threads that fill them instead of being
// zero-filled and page-faulted by one thread
template <class T>
struct fvm_noinit_alloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = fvm_noinit_alloc<U>;
    };
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(static_cast<A&&>(a)...);
    }
};
template <class T>
using fvm_rawvec = std::vector<T, fvm_noinit_alloc<T>>;

template <class Tp, class Al>
int32_t fvm_dev_upload(fvm_ctx* h, Tp** p, const std::vector<Tp, Al>& v) {
    int32_t rc = fvm_dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty())
        FVM_CUDA(h, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(Tp), cudaMemcpyHostToDevice, h->stream));
    return FVM_OK;
}

// implemented in fvm_rhs.cu
int32_t fvm_launch_geometry(fvm_ctx* h, const int32_t* d_tri_native /*[T][3] native node ids*/);
int32_t fvm_launch_volumes(fvm_ctx* h);
int32_t fvm_launch_rhs(fvm_ctx* h, double t, const double* u, double* du);
int32_t fvm_launch_dirichlet(fvm_ctx* h, double t, double* u);
int32_t fvm_launch_permute(fvm_ctx* h, const double* src, double* dst, bool to_native);
int32_t fvm_export_geometry(fvm_ctx* h, double* s9, double* mid6, double* nrm6, double* len3);  // device, native tri order
int32_t fvm_ensure_state(fvm_ctx* h);
void fvm_shard_release(fvm_ctx* h);
int32_t fvm_halo_exchange(fvm_ctx* h, double* u_native);
int32_t fvm_halo_begin(fvm_ctx* h, double* u_native);
int32_t fvm_halo_done(fvm_ctx* h);  // marks the end of the work queued on the communication stream
int32_t fvm_halo_wait(fvm_ctx* h);
// operator applications with the ghost refresh (overlapped with the independent tiles when possible)
int32_t fvm_apply_rhs(fvm_ctx* h, double t, double* x, double* out);
int32_t fvm_apply_spmv(fvm_ctx* h, double* x, double* out, bool add_b, bool scale);
// part: 0 = everything, 1 = independent tiles only, 2 = halo-dependent tiles (+ boundary-edge kernel),
// 3 = the kernels that need every tile (interface / tail rows)
//       4 = the tiles h->pipe_list[pipe_off .. pipe_off + pipe_count) only
int32_t fvm_launch_rhs_part(fvm_ctx* h, double t, const double* u, double* du, int part);
// streaming recompute kernel over tiles list[off .. off+count) (list == nullptr: tiles 0 .. count)
int32_t fvm_launch_rhs_stream(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count);
int32_t fvm_stream_fill_vinv(fvm_ctx* h);  // after the control volumes exist: 1/V of every interior node into the packs
// interface nodes list[off .. off+count) (indices into ifc_node); points that are not vertices (du = 0)
int32_t fvm_launch_rhs_interface_list(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count);
int32_t fvm_launch_rhs_nonvertex(fvm_ctx* h, double* du);
// live boundary edges list[off .. off+count) (indices into the live-edge records) -> partial slots
int32_t fvm_launch_rhs_boundary_list(fvm_ctx* h, double t, const double* u, const int32_t* list, int off, int count);
int32_t fvm_rhs_pipelined(fvm_ctx* h, double t, const double* u_host, double* du_host, bool* used);
int32_t fvm_spmv_pipelined(fvm_ctx* h, const double* x_host, double* y_host, bool add_b, bool* used);
int32_t fvm_launch_spmv_tail_list(fvm_ctx* h, const double* x, double* y, bool add_b, const int32_t* list, int off, int count);
void fvm_pipe_release(fvm_ctx* h);
void fvm_pipe_report(fvm_ctx* h, int mode, double seconds);
int32_t fvm_launch_spmv_part(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale, int part);
// Work fused into the tile / tail SpMV kernels by the Krylov solvers (fvm_solvers.cu):
//   kind 2: sum_i x[i] * y[i] over the CTA's rows goes to dotpart[tile] / dotpart[n_tiles + blockIdx.x] (fixed-shape
//           block reduction: deterministic) -- the p.(A p) of CG without a separate pass over p and q; nothing runs
//           once sc[SC_DONE] is set.  Works sharded as well (ghost rows of A are zero: they add nothing).
// Measured and NOT kept (r2q/r2r/r2s logs): forming the input (p = z + beta p, or a Tsit5 stage combination)
// while x is staged, or the next stage input in the epilogue -- the tile kernel is latency-bound in its staging phase
// and its epilogue, so the extra loads cost more there (64 -> 116 us at 2048^2) than the streaming kernels they replace.
struct SpmvFuse {
    int32_t kind = 0;
    const double* sc = nullptr;
    double* dotpart = nullptr;
};
bool fvm_spmv_fusable(fvm_ctx* h);            // the tile / sliced-ELL kernels are in use
int32_t fvm_spmv_fused_partials(fvm_ctx* h);  // entries of dotpart a kind-2 application writes (tiles + tail CTAs)
int32_t fvm_apply_spmv_fused(fvm_ctx* h, double* x, double* out, bool add_b, bool scale, const SpmvFuse& f);
// device scalars of the Krylov solvers / adaptive stepper (d_red + 8 * 2048)
enum { SC_RZ = 0, SC_PQ, SC_ALPHA, SC_BETA, SC_RR, SC_BNORM2, SC_RHO, SC_OMEGA, SC_TS, SC_TT, SC_RHV, SC_DONE, SC_ITER, SC_TOL2, SC_RESTART, SC_SUM0, SC_SUM1, SC_SUM2, SC_RZ2, SC_N };
int32_t fvm_allreduce_sum(fvm_ctx* h, double* d_vals, int n);
int32_t fvm_global_or(fvm_ctx* h, bool local, bool* global);
#define FVM_NODE_GHOST 4
int32_t fvm_launch_spmv(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale);
int32_t fvm_launch_jacobian(fvm_ctx* h, double t, const double* u_native);  // fvm_jacobian.cu: block values -> h->jac_val
void fvm_prof_begin(fvm_ctx* h);
void fvm_prof_end(fvm_ctx* h);
