// libfvmcuda: fvm_rhs with HOST buffers as a three-stream pipeline.
//
// The drop-in call `fvm_eqs!(du, u, p, t)` (/root/reference/src/equations/main_equations.jl:28-35) hands the
// library host vectors, so the call is bound by PCIe: 8*neq*N bytes in, 8*neq*N bytes out, one after the
// other around ~0.7 ms of kernels.  PCIe is full duplex, so the vector is cut into K bands of consecutive
// CALLER indices and the three steps overlap:
//
//   copy-in stream : the K bands H2D, back to back                                          -> event in[b]
//   compute stream : stage s waits for in[s], scatters band s into native (tile-major) order, runs the tiles
//                    and live boundary edges whose nodes all lie in bands <= s, then the interface nodes
//                    whose tiles / edges have all run, then gathers the output bands that are now final
//                    back to caller order                                                   -> event stage[s]
//   copy-out stream: D2H of those bands, while later bands are still arriving.
//
// Which stage a tile / interface node / output band belongs to is a property of the mesh numbering and is
// computed once per handle (lazily, at the first host-buffer call).  On a lattice in the reference's
// row-major numbering a Hilbert tile spans ~2 sqrt(TT/2) rows, so nearly every tile is ready one band after
// its rows arrive; for an arbitrary numbering the plan degenerates gracefully to "everything in the last
// stage", i.e. the unpipelined schedule.  Every node is still summed by the same kernels in the same order,
// so the result is bit-identical to the unpipelined path.
#include <algorithm>

#include <omp.h>

#include "fvm_internal.h"

#ifndef FVM_PIPE_ZC_DEFAULT
#define FVM_PIPE_ZC_DEFAULT 0
#endif

namespace {

struct PipePlan {
    int K = 0;
    int mode = 0;  // 0: fvm_rhs (tiles, boundary edges, interface nodes)   1: fvm_spmv (tiles, tail-row slices)
    bool useful = false;
    std::vector<int64_t> band_lo;         // [K+1] caller node index boundaries
    std::vector<int32_t> tile_stage_ptr;  // [K+1] into d_tile_order
    std::vector<int32_t> ifc_stage_ptr;   // [K+1] into d_ifc_order
    std::vector<int32_t> edge_stage_ptr;  // [K+1] into d_edge_order (live boundary edges)
    std::vector<int64_t> out_lo;          // [K+1] stage s completes the caller indices out_lo[s] <= j < out_lo[s+1]
    int64_t head = 0;                     // ... and the last stage also [0, head)  (out_lo[0] == head)
    int32_t* d_tile_order = nullptr;
    int32_t* d_ifc_order = nullptr;
    int32_t* d_edge_order = nullptr;
    double* d_out = nullptr;              // caller-order staging of du
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_stage;
    cudaEvent_t ev_start = nullptr, ev_out = nullptr;
    // one-off choice between the banded pipeline and the plain schedule (H2D, kernels, D2H in sequence): call 0 warms
    // the pipeline up, call 1 times it, call 2 times the plain schedule, later calls take the faster one.  Several
    // ranks sharing one host can saturate its memory system, where the three concurrent streams of the pipeline lose
    // (measured at 8 ranks: 16.3 ms pipelined vs 15.0 ms plain).  Both schedules give bit-identical results and contain
    // exactly one halo exchange, so ranks may choose differently.
    int calls = 0;
    double t_pipe = 0.0, t_plain = 0.0;
};

__global__ void band_scatter_kernel(const int32_t* __restrict__ new_of_old, const double* __restrict__ src_caller,
                                    double* __restrict__ dst_native, const int64_t lo, const int64_t hi, const int neq) {
    const int64_t idx = lo * neq + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hi * neq) return;
    const int64_t j = idx / neq;
    const int v = (int)(idx - j * neq);
    dst_native[(int64_t)new_of_old[j] * neq + v] = src_caller[idx];
}

__global__ void band_gather_kernel(const int32_t* __restrict__ new_of_old, const double* __restrict__ src_native,
                                   double* __restrict__ dst_caller, const int64_t lo, const int64_t hi, const int neq) {
    const int64_t idx = lo * neq + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hi * neq) return;
    const int64_t j = idx / neq;
    const int v = (int)(idx - j * neq);
    dst_caller[idx] = src_native[(int64_t)new_of_old[j] * neq + v];
}

// Zero-copy forms of the two kernels above: the caller's page-locked HOST vector is read / written through its device
// alias, so a band crosses PCIe and changes numbering in ONE kernel -- no copy-engine launch per band (each
// cudaMemcpyAsync + event costs the link ~15-20 us of idle time: 20 copies per call) and no staging pass through HBM.
// A few persistent CTAs keep enough 8-byte loads in flight to fill the link (4 per thread before the first store).
__global__ void __launch_bounds__(256) band_scatter_host_kernel(const int32_t* __restrict__ new_of_old, const double* __restrict__ src_host,
                                                                double* __restrict__ dst_native, const int64_t lo, const int64_t hi, const int neq) {
    const int64_t n = (hi - lo) * neq, base = lo * neq, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * stride < n) v[k] = __ldcs(src_host + base + i0 + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t idx = base + i0 + k * stride;
            if (idx < base + n) {
                const int64_t j = idx / neq;
                dst_native[(int64_t)new_of_old[j] * neq + (idx - j * neq)] = v[k];
            }
        }
    }
}

__global__ void __launch_bounds__(256) band_gather_host_kernel(const int32_t* __restrict__ new_of_old, const double* __restrict__ src_native,
                                                               double* __restrict__ dst_host, const int64_t lo, const int64_t hi, const int neq) {
    const int64_t n = (hi - lo) * neq, base = lo * neq, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t idx = base + i0 + k * stride;
            if (idx < base + n) {
                const int64_t j = idx / neq;
                v[k] = src_native[(int64_t)new_of_old[j] * neq + (idx - j * neq)];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * stride < n) __stcs(dst_host + base + i0 + k * stride, v[k]);
    }
}

// device alias of a page-locked host range (cudaHostAlloc / cudaHostRegister / fvm_host_register), or null
const void* host_alias(const void* p, size_t bytes) {
    cudaPointerAttributes a0, a1;
    if (!p || bytes == 0 || cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, (const char*)p + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
    if ((const char*)a1.devicePointer - (const char*)a0.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;
    return a0.devicePointer;
}

}  // namespace

static void release_plan(void*& slot) {
    PipePlan* P = (PipePlan*)slot;
    if (!P) return;
    for (cudaEvent_t e : P->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : P->ev_stage) cudaEventDestroy(e);
    if (P->ev_start) cudaEventDestroy(P->ev_start);
    if (P->ev_out) cudaEventDestroy(P->ev_out);
    if (P->s_in) cudaStreamDestroy(P->s_in);
    if (P->s_out) cudaStreamDestroy(P->s_out);
    delete P;  // the device arrays are in h->allocs
    slot = nullptr;
}

void fvm_pipe_release(fvm_ctx* h) {
    release_plan(h->pipe);
    release_plan(h->pipe_spmv);
}

static int32_t build_plan(fvm_ctx* h, int K, int mode, void*& slot) {
    PipePlan* P = new PipePlan;
    slot = P;
    P->K = K;
    P->mode = mode;
    const int64_t N = h->N;
    const int32_t n_tiles = h->dm.n_tiles, n_ifc = h->dm.n_ifc, n_vertices = h->dm.n_vertices;
    // Band borders.  The first output band can leave only after the second stage and the last two output bands
    // only after the last input band, so the call costs about (first two bands) + max(copy-in, copy-out) of the
    // rest + (last two bands): the first and last two bands get half the width of the others (FVM_PIPE_TAPER=0:
    // uniform bands).
    P->band_lo.assign(K + 1, 0);
    {
        const char* e = getenv("FVM_PIPE_TAPER");
        const int taper = K < 6 ? 0 : (e ? atoi(e) : 1);  // 0: uniform bands, 1: first / last two at half width, 2: quarter, half, ...
        std::vector<int64_t> w(K, 4);
        if (taper == 1) w[0] = w[1] = w[K - 2] = w[K - 1] = 2;
        if (taper >= 2 && K >= 8) {
            w[0] = w[1] = w[K - 2] = w[K - 1] = 1;
            w[2] = w[K - 3] = 2;
        }
        int64_t tot = 0, acc = 0;
        for (int b = 0; b < K; ++b) tot += w[b];
        for (int b = 0; b < K; ++b) {
            acc += w[b];
            P->band_lo[b + 1] = N * acc / tot;
        }
    }
    // node j is in band b iff band_lo[b] <= j < band_lo[b+1]
    auto band = [&](int64_t j) { return (int)(std::upper_bound(P->band_lo.begin() + 1, P->band_lo.end() - 1, j) - (P->band_lo.begin() + 1)); };
    const int32_t* old_of_new = h->node_old_of_new.data();
    // ---- stage of every tile: the last band that holds one of its (own or external) nodes -------------
    std::vector<int32_t> tile_stage(n_tiles, 0);
#pragma omp parallel for schedule(static)
    for (int32_t t = 0; t < n_tiles; ++t) {
        int s = 0;
        for (int32_t g = h->h_tile_node0[t]; g < h->h_tile_node0[t] + h->h_tile_nown[t]; ++g) s = std::max(s, band(old_of_new[g]));
        for (int32_t k = h->h_tile_ext0[t]; k < h->h_tile_ext0[t + 1]; ++k) s = std::max(s, band(old_of_new[h->h_ext_ids[k]]));
        tile_stage[t] = s;
    }
    // sharded: ghost entries are refreshed by the halo exchange, which needs the packed owned values of every
    // band and therefore runs at the head of the last stage; whatever reads a ghost node runs after it
    const bool sharded = h->halo_ready && !h->h_ghost.empty();
    auto is_ghost = [&](int32_t g) { return sharded && h->h_ghost[old_of_new[g]] != 0; };
    if (sharded) {
#pragma omp parallel for schedule(static)
        for (int32_t t = 0; t < n_tiles; ++t) {
            bool d = false;
            for (int32_t g = h->h_tile_node0[t]; !d && g < h->h_tile_node0[t] + h->h_tile_nown[t]; ++g) d = is_ghost(g);
            for (int32_t k = h->h_tile_ext0[t]; !d && k < h->h_tile_ext0[t + 1]; ++k) d = is_ghost(h->h_ext_ids[k]);
            if (d) tile_stage[t] = K - 1;
        }
    }
    // ---- stage of every live boundary edge: the three vertices of its triangle have arrived -----------
    const int32_t n_edges = mode == 0 ? (int32_t)h->h_bnd.size() : 0;  // the operator has no edge pass
    std::vector<int32_t> edge_stage(n_edges, 0);
    for (int32_t e = 0; e < n_edges; ++e)
        for (int q = 0; q < 3; ++q)
            edge_stage[e] = std::max(edge_stage[e], is_ghost(h->h_bnd[e].v[q]) ? K - 1 : band(old_of_new[h->h_bnd[e].v[q]]));
    // ---- stage of every interface node: all tiles and boundary edges that feed it have run -----------
    std::vector<int32_t> ifc_stage(n_ifc, 0);
    {
        int32_t base = 0;  // interface nodes are numbered tile by tile (fvm_finalize step 4)
        std::vector<int32_t> ifc_base(n_tiles + 1, 0);
        for (int32_t t = 0; t < n_tiles; ++t) {
            ifc_base[t] = base;
            base += h->h_tile_nown[t] - h->h_tile_nint[t];
        }
        for (int32_t t = 0; t < n_tiles; ++t) {
            for (int32_t q = 0; q < h->h_tile_nown[t] - h->h_tile_nint[t]; ++q)
                ifc_stage[ifc_base[t] + q] = std::max(ifc_stage[ifc_base[t] + q], tile_stage[t]);
            for (int32_t k = h->h_tile_ext0[t]; k < h->h_tile_ext0[t + 1]; ++k) {
                const int32_t g = h->h_ext_ids[k];
                const int32_t i = (int32_t)(std::lower_bound(h->h_ifc_node.begin(), h->h_ifc_node.end(), g) - h->h_ifc_node.begin());
                if (i >= n_ifc || h->h_ifc_node[i] != g) return fvm_fail(h, FVM_ERR_STATE, "pipeline plan: external node is not an interface node");
                ifc_stage[i] = std::max(ifc_stage[i], tile_stage[t]);
            }
        }
        for (int32_t e = 0; e < n_edges; ++e)
            for (int q = 0; q < 2; ++q) {
                const int32_t g = h->h_bnd[e].v[q == 0 ? h->h_bnd[e].pi : h->h_bnd[e].pj];
                const int32_t i = (int32_t)(std::lower_bound(h->h_ifc_node.begin(), h->h_ifc_node.end(), g) - h->h_ifc_node.begin());
                if (i >= n_ifc || h->h_ifc_node[i] != g) return fvm_fail(h, FVM_ERR_STATE, "pipeline plan: boundary-edge endpoint is not an interface node");
                ifc_stage[i] = std::max(ifc_stage[i], edge_stage[e]);
            }
    }
    // ---- counting sorts -> launch lists ------------------------------------------------------------------
    auto sort_by_stage = [&](const std::vector<int32_t>& stage, std::vector<int32_t>& ptr, std::vector<int32_t>& order) {
        ptr.assign(K + 1, 0);
        for (int32_t s : stage) ptr[s + 1]++;
        for (int s = 0; s < K; ++s) ptr[s + 1] += ptr[s];
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
        order.resize(stage.size());
        for (size_t i = 0; i < stage.size(); ++i) order[fill[stage[i]]++] = (int32_t)i;
    };
    // fvm_spmv: the interface rows (then the points that are not vertices) form the tail rows, 32 per sliced-ELL
    // slice; the columns of an interface row are vertices of triangles that hold the node, i.e. local nodes of
    // the tiles counted in ifc_stage, so the row can run in that stage; a slice runs when its 32 rows can
    std::vector<int32_t> unit_stage = ifc_stage;  // what the third kernel of a stage iterates over
    if (mode == 1) {
        const int32_t n_tail = n_ifc + (int32_t)(N - n_vertices), n_slices = (n_tail + 31) / 32;
        unit_stage.assign(n_slices, 0);
        for (int32_t k = 0; k < n_tail; ++k) unit_stage[k / 32] = std::max(unit_stage[k / 32], k < n_ifc ? ifc_stage[k] : K - 1);
        for (int32_t i = 0; i < n_ifc; ++i) ifc_stage[i] = unit_stage[i / 32];  // when the row's result is final
    }
    std::vector<int32_t> tile_order, ifc_order, edge_order;
    sort_by_stage(tile_stage, P->tile_stage_ptr, tile_order);
    sort_by_stage(unit_stage, P->ifc_stage_ptr, ifc_order);
    sort_by_stage(edge_stage, P->edge_stage_ptr, edge_order);
    // ---- output ranges -----------------------------------------------------------------------------------
    // du[j] is final after stage fs(j) (its tile for an interior node, the last tile / edge feeding it for an interface
    // node).  Stage s sends the caller indices [out_lo[s], out_lo[s+1]), where out_lo[s+1] is the longest PREFIX of the
    // vector that is final after stage s: contiguous ranges whatever the numbering.  On a row-major lattice that prefix
    // ends one tile height (~2 sqrt(TT/2) rows) before the end of input band s, so an output range leaves right after the
    // stage of "its" input band (round 1 sent whole input bands, i.e. one stage -- a full band -- later, and the call ended
    // with two bands of copy-out after the last copy-in instead of one: profiles/r02_e2e_timeline.txt).
    // A sharded strip also has a HEAD that is final only after the last stage (the rows next to the lower neighbour wait for
    // the halo exchange): whatever of the first input band ends that late is sent as a second range of the last stage and
    // the prefix rule applies behind it.
    const int32_t* new_of_old = h->node_new_of_old.data();
    std::vector<uint8_t> fs(N);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const int32_t g = new_of_old[j];
        int s;
        if (g >= n_vertices) {
            s = K - 1;  // points that are not vertices are zeroed in the last stage
        } else {
            const int32_t t = (int32_t)(std::upper_bound(h->h_tile_node0.begin(), h->h_tile_node0.end(), g) - h->h_tile_node0.begin()) - 1;
            int32_t tt = t;  // tiles without own nodes share node0 with their successor: step back to the owner
            while (tt > 0 && g >= h->h_tile_node0[tt] + h->h_tile_nown[tt]) --tt;
            const int32_t l = g - h->h_tile_node0[tt];
            if (l < h->h_tile_nint[tt]) {
                s = tile_stage[tt];
            } else {
                const int32_t i = (int32_t)(std::lower_bound(h->h_ifc_node.begin(), h->h_ifc_node.end(), g) - h->h_ifc_node.begin());
                s = ifc_stage[i];
            }
        }
        fs[j] = (uint8_t)s;
    }
    int64_t head = 0;
    if (K > 1 && fs[0] == K - 1)  // (a vector whose first entry is early has no late head)
        for (int64_t j = P->band_lo[1] - 1; j >= 0; --j)
            if (fs[j] == K - 1) {
                head = j + 1;
                break;
            }
    std::vector<int64_t> first_of_stage(K, N);  // smallest caller index >= head whose result is final only after stage f
    {
        const int nth = omp_get_max_threads();
        std::vector<int64_t> loc((size_t)nth * K, N);
#pragma omp parallel
        {
            int64_t* mine = loc.data() + (size_t)omp_get_thread_num() * K;
#pragma omp for schedule(static)
            for (int64_t j = head; j < N; ++j)
                if (j < mine[fs[j]]) mine[fs[j]] = j;
        }
        for (int th = 0; th < nth; ++th)
            for (int f = 0; f < K; ++f) first_of_stage[f] = std::min(first_of_stage[f], loc[(size_t)th * K + f]);
    }
    P->out_lo.assign(K + 1, head);
    P->out_lo[K] = N;
    for (int s = K - 2; s >= 0; --s) P->out_lo[s + 1] = std::max(head, std::min(P->out_lo[s + 2], first_of_stage[s + 1]));
    P->head = head;
    // worth it only if a good part of the output leaves before the last stage
    int early = 0;
    for (int s = 0; s + 1 < K; ++s) early += P->out_lo[s + 1] > P->out_lo[s];
    P->useful = 2 * (P->out_lo[K - 1] - head) >= N;
    int32_t rc;
    if ((rc = fvm_dev_upload(h, &P->d_tile_order, tile_order))) return rc;
    if ((rc = fvm_dev_upload(h, &P->d_ifc_order, ifc_order))) return rc;
    if ((rc = fvm_dev_upload(h, &P->d_edge_order, edge_order))) return rc;
    if ((rc = fvm_dev_alloc(h, &P->d_out, (size_t)N * h->neq))) return rc;
    FVM_CUDA(h, cudaStreamCreateWithFlags(&P->s_in, cudaStreamNonBlocking));
    FVM_CUDA(h, cudaStreamCreateWithFlags(&P->s_out, cudaStreamNonBlocking));
    P->ev_in.resize(K);
    P->ev_stage.resize(K);
    for (int b = 0; b < K; ++b) {
        FVM_CUDA(h, cudaEventCreateWithFlags(&P->ev_in[b], cudaEventDisableTiming));
        FVM_CUDA(h, cudaEventCreateWithFlags(&P->ev_stage[b], cudaEventDisableTiming));
    }
    FVM_CUDA(h, cudaEventCreateWithFlags(&P->ev_start, cudaEventDisableTiming));
    FVM_CUDA(h, cudaEventCreateWithFlags(&P->ev_out, cudaEventDisableTiming));
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));  // the uploads above
    h->stats[12] = K;
    h->stats[13] = early;
    return FVM_OK;
}

// the schedule shared by fvm_rhs and fvm_spmv; stage(s) queues the kernels of stage s on the compute stream
template <class StageFn>
static int32_t run_pipeline(fvm_ctx* h, PipePlan& P, const double* in_host, double* out_host, StageFn stage) {
    const int K = P.K, neq = h->neq;
    cudaStream_t sc = h->stream;
    // FVM_PIPE_TRACE=1: device timestamps of every band's copy-in, stage and copy-out on stderr (diagnostics only)
    static const bool trace = getenv("FVM_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tr;  // [0]: start, then K x (in, stage, out)
    if (trace) {
        tr.resize(1 + 3 * K);
        for (cudaEvent_t& e : tr) cudaEventCreate(&e);
    }
    // The copy streams carry nothing but copies, back to back, so that neither DMA engine ever waits for a
    // kernel: the scatter of band s and the gathers of the bands that stage s completes run on the compute
    // stream, which is idle most of the time (a stage is ~0.09 ms of kernels per ~0.3 ms of copy at 16.7M nodes).
    // FVM_PIPE_ZC: bit 0 = copy-in by kernel, bit 1 = copy-out by kernel (page-locked buffers only)
    const char* e_zc = getenv("FVM_PIPE_ZC");
    const char* e_ctas = getenv("FVM_PIPE_ZC_CTAS");
    const int zc_env = e_zc ? atoi(e_zc) : FVM_PIPE_ZC_DEFAULT;
    const int zc_ctas = e_ctas ? std::max(1, atoi(e_ctas)) : 64;
    const double* in_dev = (zc_env & 1) ? (const double*)host_alias(in_host, sizeof(double) * (size_t)h->N * neq) : nullptr;
    double* out_dev = (zc_env & 2) ? (double*)host_alias(out_host, sizeof(double) * (size_t)h->N * neq) : nullptr;
    FVM_CUDA(h, cudaEventRecord(P.ev_start, sc));  // an earlier call's kernels precede the first copy-in
    if (trace) cudaEventRecord(tr[0], sc);
    FVM_CUDA(h, cudaStreamWaitEvent(P.s_in, P.ev_start, 0));
    FVM_CUDA(h, cudaStreamWaitEvent(P.s_out, P.ev_start, 0));
    for (int b = 0; b < K; ++b) {
        const int64_t lo = P.band_lo[b], hi = P.band_lo[b + 1], cnt = (hi - lo) * neq;
        if (cnt > 0 && in_dev)
            band_scatter_host_kernel<<<zc_ctas, 256, 0, P.s_in>>>(h->d_node_new_of_old, in_dev, h->d_u, lo, hi, neq);
        else if (cnt > 0)
            FVM_CUDA(h, cudaMemcpyAsync(h->d_io + lo * neq, in_host + lo * neq, sizeof(double) * cnt, cudaMemcpyHostToDevice, P.s_in));
        FVM_CUDA(h, cudaEventRecord(P.ev_in[b], P.s_in));
        if (trace) cudaEventRecord(tr[1 + 3 * b], P.s_in);
    }
    int32_t rc = FVM_OK;
    for (int s = 0; s < K && !rc; ++s) {
        FVM_CUDA(h, cudaStreamWaitEvent(sc, P.ev_in[s], 0));
        {
            const int64_t lo = P.band_lo[s], hi = P.band_lo[s + 1], cnt = (hi - lo) * neq;
            if (cnt > 0 && !in_dev)
                band_scatter_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, sc>>>(h->d_node_new_of_old, h->d_io, h->d_u, lo, hi, neq);
        }
        if ((rc = stage(s))) break;
        if (trace) cudaEventRecord(tr[2 + 3 * s], sc);
        const int64_t rng[2][2] = {{P.out_lo[s], P.out_lo[s + 1]}, {0, s == K - 1 ? P.head : 0}};
        if (rng[0][1] <= rng[0][0] && rng[1][1] <= rng[1][0]) continue;
        if (!out_dev)
            for (const auto& r : rng)
                if (r[1] > r[0])
                    band_gather_kernel<<<(unsigned)(((r[1] - r[0]) * neq + 255) / 256), 256, 0, sc>>>(h->d_node_new_of_old, h->d_du, P.d_out, r[0], r[1], neq);
        FVM_CUDA(h, cudaEventRecord(P.ev_stage[s], sc));
        FVM_CUDA(h, cudaStreamWaitEvent(P.s_out, P.ev_stage[s], 0));
        for (const auto& r : rng) {
            if (r[1] <= r[0]) continue;
            if (out_dev)
                band_gather_host_kernel<<<zc_ctas, 256, 0, P.s_out>>>(h->d_node_new_of_old, h->d_du, out_dev, r[0], r[1], neq);
            else
                FVM_CUDA(h, cudaMemcpyAsync(out_host + r[0] * neq, P.d_out + r[0] * neq, sizeof(double) * (r[1] - r[0]) * neq, cudaMemcpyDeviceToHost, P.s_out));
        }
        if (trace) cudaEventRecord(tr[3 + 3 * s], P.s_out);
    }
    // leave the three streams joined whatever happened, so that the handle stays usable after an error
    cudaEventRecord(P.ev_out, P.s_out);
    cudaStreamWaitEvent(sc, P.ev_out, 0);
    cudaStreamWaitEvent(sc, P.ev_in[K - 1], 0);
    cudaError_t ce = cudaStreamSynchronize(sc);
    if (trace) {
        if (!rc && ce == cudaSuccess)
            for (int b = 0; b < K; ++b) {
                float ti = 0, ts = 0, to = 0;
                cudaEventElapsedTime(&ti, tr[0], tr[1 + 3 * b]);
                cudaEventElapsedTime(&ts, tr[0], tr[2 + 3 * b]);
                cudaEventElapsedTime(&to, tr[0], tr[3 + 3 * b]);
                fprintf(stderr, "[fvm_pipe] stage %2d  nodes in %9lld  copy-in done %7.3f  kernels done %7.3f   nodes out %9lld  copy-out done %7.3f ms\n", b,
                        (long long)(P.band_lo[b + 1] - P.band_lo[b]), ti, ts, (long long)(P.out_lo[b + 1] - P.out_lo[b]), to);
            }
        for (cudaEvent_t e : tr) cudaEventDestroy(e);
        cudaGetLastError();
    }
    if (rc) return rc;
    FVM_CUDA(h, ce);
    FVM_CUDA(h, cudaGetLastError());
    h->stats[14] += 1;
    return FVM_OK;
}

// common gate: returns the plan to use, or null for the plain schedule
static int32_t pipeline_plan(fvm_ctx* h, int mode, void*& slot, PipePlan** out) {
    *out = nullptr;
    const char* e_off = getenv("FVM_NO_PIPELINE");
    if ((e_off && e_off[0] == '1') || h->profiling) return FVM_OK;
    if (h->nranks > 1 && !h->halo_ready) return FVM_OK;  // sharded handle whose halo plan is not installed yet
    const char* e_min = getenv("FVM_PIPE_MIN_NODES");
    const int64_t min_nodes = e_min ? atoll(e_min) : (int64_t)1 << 20;
    if (h->N < min_nodes) return FVM_OK;
    if (!slot) {
        const char* e_k = getenv("FVM_PIPE_BANDS");
        int K = e_k ? atoi(e_k) : 10;
        K = (int)std::max<int64_t>(2, std::min<int64_t>(std::min(K, 64), h->N));
        int32_t rc = build_plan(h, K, mode, slot);
        if (rc) return rc;
    }
    PipePlan* P = (PipePlan*)slot;
    const char* e_force = getenv("FVM_PIPE_FORCE");  // tests: run the pipeline even where it cannot overlap anything
    const bool force = e_force && e_force[0] == '1';
    if (!P->useful && !force) return FVM_OK;
    static const bool tune = !(getenv("FVM_PIPE_AUTOTUNE") && getenv("FVM_PIPE_AUTOTUNE")[0] == '0');
    if (tune && !force) {
        if (P->calls == 2) return FVM_OK;                               // the timed plain call
        if (P->calls > 2 && P->t_plain < P->t_pipe) return FVM_OK;      // plain won
    }
    *out = P;
    return FVM_OK;
}

// wall time of one host-buffer call, reported by fvm_rhs / fvm_spmv (mode 0 / 1) for the one-off schedule choice
void fvm_pipe_report(fvm_ctx* h, int mode, double seconds) {
    PipePlan* P = (PipePlan*)(mode == 0 ? h->pipe : h->pipe_spmv);
    if (!P) return;
    if (P->calls == 1) P->t_pipe = seconds;
    else if (P->calls == 2) P->t_plain = seconds;
    if (P->calls == 2) h->pipe_choice[mode] = P->t_plain < P->t_pipe ? 2 : 1;  // 1: pipeline kept, 2: plain schedule chosen
    if (P->calls < 3) P->calls += 1;
}

int32_t fvm_rhs_pipelined(fvm_ctx* h, double t, const double* u_host, double* du_host, bool* used) {
    *used = false;
    if (u_host == du_host) return FVM_OK;
    PipePlan* Pp = nullptr;
    int32_t rc = pipeline_plan(h, 0, h->pipe, &Pp);
    if (rc || !Pp) return rc;
    PipePlan& P = *Pp;
    const int K = P.K;
    rc = run_pipeline(h, P, u_host, du_host, [&](int s) -> int32_t {
        int32_t r;
        if (s == K - 1 && (r = fvm_halo_exchange(h, h->d_u))) return r;  // no-op unless sharded
        r = fvm_launch_rhs_boundary_list(h, t, h->d_u, P.d_edge_order, P.edge_stage_ptr[s], P.edge_stage_ptr[s + 1] - P.edge_stage_ptr[s]);
        if (r) return r;
        h->pipe_list = P.d_tile_order;
        h->pipe_off = P.tile_stage_ptr[s];
        h->pipe_count = P.tile_stage_ptr[s + 1] - P.tile_stage_ptr[s];
        if (h->pipe_count > 0 && (r = fvm_launch_rhs_part(h, t, h->d_u, h->d_du, 4))) return r;
        if ((r = fvm_launch_rhs_interface_list(h, t, h->d_u, h->d_du, P.d_ifc_order, P.ifc_stage_ptr[s], P.ifc_stage_ptr[s + 1] - P.ifc_stage_ptr[s])))
            return r;
        return s == K - 1 ? fvm_launch_rhs_nonvertex(h, h->d_du) : FVM_OK;
    });
    if (rc) return rc;
    *used = true;
    return FVM_OK;
}

// y = A x (+ b) with host vectors: tiles own their interior rows; interface rows and points that are not
// vertices are the sliced-ELL tail rows, run slice by slice as their columns arrive
int32_t fvm_spmv_pipelined(fvm_ctx* h, const double* x_host, double* y_host, bool add_b, bool* used) {
    *used = false;
    if (x_host == y_host || !h->csr.use_tile_spmv || h->neq != 1) return FVM_OK;
    PipePlan* Pp = nullptr;
    int32_t rc = pipeline_plan(h, 1, h->pipe_spmv, &Pp);
    if (rc || !Pp) return rc;
    PipePlan& P = *Pp;
    rc = run_pipeline(h, P, x_host, y_host, [&](int s) -> int32_t {
        int32_t r;
        if (s == P.K - 1 && (r = fvm_halo_exchange(h, h->d_u))) return r;  // no-op unless sharded
        h->pipe_list = P.d_tile_order;
        h->pipe_off = P.tile_stage_ptr[s];
        h->pipe_count = P.tile_stage_ptr[s + 1] - P.tile_stage_ptr[s];
        if (h->pipe_count > 0 && (r = fvm_launch_spmv_part(h, h->d_u, h->d_du, add_b, false, 4))) return r;
        return fvm_launch_spmv_tail_list(h, h->d_u, h->d_du, add_b, P.d_ifc_order, P.ifc_stage_ptr[s], P.ifc_stage_ptr[s + 1] - P.ifc_stage_ptr[s]);
    });
    if (rc) return rc;
    *used = true;
    return FVM_OK;
}
