// libfvmcuda: the streaming recompute-geometry RHS kernel (geometry_mode = 1).
//
//   du = fvm_eqs!(du, u, p, t)      /root/reference/src/equations/main_equations.jl:38-44
//
// Same mathematics and the same per-node summation order as rhs_tile_kernel (fvm_rhs.cu), but built as a
// persistent, warp-specialised pipeline so that HBM, the fp64 pipe and the issue slots work at the same time:
//
//   * a CTA walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  (grid = SMs x resident CTAs);
//   * one PRODUCER warp runs ahead: per tile it issues ONE bulk async copy (cp.async.bulk, the TMA engine) of the
//     tile's static pack (TilePackHdr: local vertex ids, vertex coordinates, 1/V, kinds, partial slots, gather
//     rows), one bulk copy of the tile's own `u` range and 8-byte cp.async gathers of the external nodes' `u`,
//     all completing on the stage's `full` mbarrier; two stages, so tile k+1 lands while tile k is computed;
//   * the CONSUMER warps never touch global memory for inputs: triangle pass (one thread per triangle:
//     geometry.jl:107-161 recomputed from the staged coordinates, shape functions, three control-volume-edge
//     fluxes, triangle_contributions.jl:28-35) -> contribution planes in shared memory -> one named barrier ->
//     node pass (a warp takes units of 32 consecutive local nodes and walks their gather rows:
//     coalesced 16-bit loads of byte offsets into the planes; the slots are
//     edge-coloured at setup so that neither the scatter nor the gather has a bank conflict), node pass of
//     source_contributions.jl:33-68 for interior nodes (coalesced `du` stores), one partial per interface node;
//   * the contribution planes are double buffered, so one barrier per tile is enough.
//
// No atomics; every output word has one writer and a fixed summation order.
#include <algorithm>

#include "fvm_device.cuh"

namespace {

constexpr int SK_STAGES = 2;
constexpr int SK_PROD = 32;  // producer threads (one warp)

struct StreamArgs {
    const uint8_t* packs;
    const int4* dir;
    const int32_t* list;  // explicit tile list (or null: tiles 0 .. count)
    int32_t list_off, count;
    int32_t pack_cap;  // bytes reserved per stage for the pack
    int32_t u_cap;     // doubles reserved per stage for the tile's u values
    int32_t prefetch;  // producer pulls the next tile's pack / u range into L2 one tile ahead
    int32_t plane;     // doubles per contribution plane: 3 * TT + 16 (the last 16 stay zero: one per bank pair)
    int32_t nbuf;      // contribution buffers: 2 (one barrier per tile) or 1 (a second barrier after the node pass, half the planes)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk async copy (TMA engine, no tensor map); 16-byte aligned addresses, size multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier gets one arrival from this thread when all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// OCC = resident CTAs per SM the register budget is planned for (shared memory decides the rest at run time)
template <int MODEL, int NEQ, int NCONS, int OCC>
__global__ void __launch_bounds__(NCONS + SK_PROD, OCC)
    rhs_stream_kernel(const DevMesh m, const FluxParams fp, const SourceParams sp, const double t, const double* __restrict__ u,
                      double* __restrict__ du, const StreamArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TT = m.tile_tris;
    const int cbuf = NEQ * a.plane;  // doubles per contribution buffer
    double* c_s = reinterpret_cast<double*>(smem_raw);
    unsigned char* stage0 = smem_raw + (size_t)a.nbuf * cbuf * sizeof(double);
    const int stage_bytes = a.pack_cap + a.u_cap * (int)sizeof(double);
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)SK_STAGES * stage_bytes);
    uint64_t* empty = full + SK_STAGES;
    const int tid = threadIdx.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SK_STAGES; ++s) {
            mbar_init(full + s, SK_PROD + 1);  // 32 cp.async arrivals + the expect_tx arrival
            mbar_init(empty + s, NCONS / 32);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < a.nbuf * NEQ * 16)  // the zero words padded gather entries point at (one per 8-byte bank pair)
        c_s[(tid / (16 * NEQ)) * cbuf + ((tid / 16) % NEQ) * a.plane + 3 * TT + (tid & 15)] = 0.0;
    __syncthreads();
    const int n_my = ((int)a.count - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid >= NCONS) {
        // ================================ producer warp ================================================
        const int lane = tid - NCONS;
        for (int i = 0; i < n_my; ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const int pos = blockIdx.x + i * gridDim.x;
            const int tile = a.list ? __ldg(a.list + a.list_off + pos) : pos;
            const int4 d0 = __ldg(a.dir + 2 * tile), d1 = __ldg(a.dir + 2 * tile + 1);
            const int node0 = d0.z, nown = d0.w, ext0 = d1.x, next = d1.y;
            mbar_wait(empty + s, ph ^ 1u);  // stage free (passes at once the first time round)
            unsigned char* st = stage0 + (size_t)s * stage_bytes;
            double* us = reinterpret_cast<double*>(st + a.pack_cap);
            const double* usrc = u + (size_t)node0 * NEQ;
            const int n = nown * NEQ;
            // bulk copies need 16-byte aligned addresses: element j of the own range sits at us[shift + j], so global
            // and shared addresses agree mod 16; an odd head / tail element goes by 8-byte cp.async
            const int shift = (int)((reinterpret_cast<uintptr_t>(usrc) >> 3) & 1);
            const int nb = n > shift ? ((n - shift) & ~1) : 0;
            if (lane == 0) {
                mbar_expect_tx(full + s, (uint32_t)d0.y + (uint32_t)nb * 8u);
                bulk_g2s(st, a.packs + ((size_t)(uint32_t)d0.x << 4), (uint32_t)d0.y, full + s);
                if (nb > 0) bulk_g2s(us + 2 * shift, usrc + shift, (uint32_t)nb * 8u, full + s);
            } else if (lane == 1) {
                if (shift && n > 0) cp_async8(us + 1, usrc);
            } else if (lane == 2) {
                if (n > shift && ((n - shift) & 1)) cp_async8(us + shift + n - 1, usrc + n - 1);
            }
            if (lane == 3 && a.prefetch && i + 1 < n_my) {
                // the pack and the u range of the NEXT tile of this CTA go to L2 now, one tile time before their bulk
                // copies are issued: those then pay an L2 hit instead of the DRAM latency
                const int npos = pos + gridDim.x;
                const int ntile = a.list ? __ldg(a.list + a.list_off + npos) : npos;
                const int4 p0 = __ldg(a.dir + 2 * ntile);
                bulk_prefetch_l2(a.packs + ((size_t)(uint32_t)p0.x << 4), (uint32_t)p0.y);
                const uintptr_t ua = (reinterpret_cast<uintptr_t>(u + (size_t)p0.z * NEQ) + 15) & ~(uintptr_t)15;
                const uintptr_t ue = reinterpret_cast<uintptr_t>(u + (size_t)(p0.z + p0.w) * NEQ) & ~(uintptr_t)15;
                if (ue > ua) bulk_prefetch_l2(reinterpret_cast<const void*>(ua), (uint32_t)(ue - ua));
            }
            for (int k = lane; k < next; k += 32) {
                const int g = __ldg(m.ext_ids + ext0 + k);
#pragma unroll
                for (int v = 0; v < NEQ; ++v) cp_async8(us + shift + (nown + k) * NEQ + v, u + (size_t)g * NEQ + v);
            }
            cp_async_arrive_noinc(full + s);
        }
        return;
    }

    // ==================================== consumer warps ================================================
    const int lane = tid & 31, warp = tid >> 5;
    constexpr bool FULL = FluxTraits<MODEL>::full;
    for (int i = 0; i < n_my; ++i) {
        const int s = i & 1;
        const uint32_t ph = (uint32_t)(i >> 1) & 1u;
        mbar_wait(full + s, ph);
        const unsigned char* st = stage0 + (size_t)s * stage_bytes;
        const int4 h0 = reinterpret_cast<const int4*>(st)[0], h1 = reinterpret_cast<const int4*>(st)[1];
        const int4 h2 = reinterpret_cast<const int4*>(st)[2], h3 = reinterpret_cast<const int4*>(st)[3];
        const int node0 = h0.x, nint = h0.y, nloc = h0.w, ntri = h1.x, nunit = h1.z;
        const ushort4* __restrict__ tri_s = reinterpret_cast<const ushort4*>(st + h2.x);
        const double2* __restrict__ xy_s = reinterpret_cast<const double2*>(st + h2.y);
        const double* __restrict__ us =
            reinterpret_cast<const double*>(st + a.pack_cap) + ((reinterpret_cast<uintptr_t>(u + (size_t)node0 * NEQ) >> 3) & 1);
        double* cb = c_s + (size_t)(a.nbuf == 2 ? (i & 1) : 0) * cbuf;

        // ---- triangle pass --------------------------------------------------------------------------
#pragma unroll 1
        for (int lt = tid; lt < ntri; lt += NCONS) {
            const ushort4 vv = tri_s[lt];
            const double2 P = xy_s[vv.x], Q = xy_s[vv.y], R = xy_s[vv.z];
            TriGeom G;
            tri_geometry<false, FULL>(P.x, P.y, Q.x, Q.y, R.x, R.y, G, nullptr);
            double dt3[3] = {0, 0, 0};
            if constexpr (FluxTraits<MODEL>::table) {
                const double* dts = reinterpret_cast<const double*>(st + h3.w);
                const int np = (ntri + 1) & ~1;
#pragma unroll
                for (int e = 0; e < 3; ++e) dt3[e] = dts[e * np + lt];
            }
            double al[NEQ], be[NEQ], ga[NEQ];  // shape_functions.jl:2-19
#pragma unroll
            for (int v = 0; v < NEQ; ++v) {
                const double ui = us[vv.x * NEQ + v], uj = us[vv.y * NEQ + v], uk = us[vv.z * NEQ + v];
                if constexpr (FULL) {  // fluxes that read u = alpha x + beta y + gamma: the reference's roundings throughout
                    al[v] = shape_coeff(G.s[0], G.s[1], G.s[2], ui, uj, uk);
                    be[v] = shape_coeff(G.s[3], G.s[4], G.s[5], ui, uj, uk);
                    ga[v] = shape_coeff(G.s[6], G.s[7], G.s[8], ui, uj, uk);
                } else {
                    al[v] = G.s[0] * ui + G.s[1] * uj + G.s[2] * uk;
                    be[v] = G.s[3] * ui + G.s[4] * uj + G.s[5] * uk;
                    ga[v] = 0.0;
                }
            }
            double Qe[3][NEQ];
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double qx[NEQ], qy[NEQ];
                flux_eval<MODEL, NEQ>(fp, FULL ? G.mx[e] : 0.0, FULL ? G.my[e] : 0.0, t, al, be, ga, dt3[e], qx, qy);
#pragma unroll
                for (int v = 0; v < NEQ; ++v) Qe[e][v] = qx[v] * G.ey[e] - qy[v] * G.ex[e];  // q . (l n),  l n = (e_y, -e_x)
            }
#pragma unroll
            for (int v = 0; v < NEQ; ++v) {  // triangle_contributions.jl:10-25; slot positions: see build_tile_packs
                double* cp = cb + v * a.plane + (lt & ~15);
                cp[vv.w & 15] = Qe[2][v] - Qe[0][v];
                cp[TT + ((vv.w >> 4) & 15)] = Qe[0][v] - Qe[1][v];
                cp[2 * TT + ((vv.w >> 8) & 15)] = Qe[1][v] - Qe[2][v];
            }
        }
        bar_sync(1, NCONS);

        // ---- node pass: a warp takes one unit of 32 consecutive local nodes at a time and walks their gather rows
        // (coalesced 16-bit codes; conflict-free plane reads by the slot colouring) -----------------------------
        const double* __restrict__ vinv_s = reinterpret_cast<const double*>(st + h2.z);
        const uint8_t* __restrict__ kind_s = st + h2.w;
        const int kstride = (nint + 15) & ~15;
        const int32_t* __restrict__ ppos_s = reinterpret_cast<const int32_t*>(st + h3.x);
        const uint16_t* __restrict__ urow = reinterpret_cast<const uint16_t*>(st + h3.y);
        const uint16_t* __restrict__ lst = reinterpret_cast<const uint16_t*>(st + h3.z);
        const unsigned char* cbytes = reinterpret_cast<const unsigned char*>(cb);
        const bool simple = (h1.y & 1) && sp.model == FVM_SRC_ZERO;
#pragma unroll 1
        // (the units left over after the last full round go to a different pair of warps every tile: with two contribution
        // buffers a warp that finishes early starts the next tile's triangle pass, so the imbalance averages out)
        for (int q = (warp + i) % (NCONS / 32); q < nunit; q += NCONS / 32) {
            const int r0 = urow[q], r1 = urow[q + 1];
            const uint16_t* lp = lst + r0 * 32 + lane;
            double acc[NEQ];
#pragma unroll
            for (int v = 0; v < NEQ; ++v) acc[v] = 0.0;
#pragma unroll 2
            for (int r = r0; r < r1; ++r, lp += 32) {
                const uint32_t code = *lp;
#pragma unroll
                for (int v = 0; v < NEQ; ++v)
                    acc[v] += *reinterpret_cast<const double*>(cbytes + code + (size_t)v * a.plane * sizeof(double));
            }
            const int l = q * 32 + lane;
            if (l < nint) {  // source_contributions.jl:33-68, finished in place
                const int g = node0 + l;
                const double vi = vinv_s[l];
                if (simple) {  // every interior node of the tile is free and there is no source term
#pragma unroll
                    for (int v = 0; v < NEQ; ++v) du[(size_t)g * NEQ + v] = acc[v] * vi;
                } else {
                    double tab[NEQ];
                    if (sp.model == FVM_SRC_TABLE) {
#pragma unroll
                        for (int v = 0; v < NEQ; ++v) tab[v] = m.src_tab[(size_t)g * NEQ + v];
                    }
#pragma unroll
                    for (int v = 0; v < NEQ; ++v) {
                        const uint8_t kind = kind_s[v * kstride + l];
                        double out;
                        if (kind == FVM_NODE_FREE) {
                            out = acc[v] * vi + source_eval<NEQ>(sp, v, us + l * NEQ, tab);
                        } else if (kind == FVM_NODE_DUDT) {
                            const CondFn c = m.cond[v * FVM_MAX_COND_FN + m.fidx[(size_t)v * m.n_nodes + g]];
                            const double2 X = xy_s[l];
                            out = cond_eval(c, X.x, X.y, t, us[l * NEQ + v]);
                        } else {
                            out = 0.0;  // Dirichlet node, ghost node
                        }
                        du[(size_t)g * NEQ + v] = out;
                    }
                }
            } else if (l < nloc) {
                const size_t p = (size_t)ppos_s[l - nint] * NEQ;
#pragma unroll
                for (int v = 0; v < NEQ; ++v) m.partial[p + v] = acc[v];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);  // this warp is done with the stage
        if (a.nbuf == 1) bar_sync(2, NCONS);    // single contribution buffer: nobody scatters into it before every node is gathered
    }
}

__global__ void pack_vinv_kernel(uint8_t* __restrict__ packs, const int4* __restrict__ dir, const double* __restrict__ vol) {
    const int4 d0 = dir[2 * blockIdx.x];
    uint8_t* base = packs + ((size_t)(uint32_t)d0.x << 4);
    const TilePackHdr H = *reinterpret_cast<const TilePackHdr*>(base);
    double* vinv = reinterpret_cast<double*>(base + H.off_vinv);
    for (int l = threadIdx.x; l < H.nint; l += blockDim.x) vinv[l] = 1.0 / vol[H.node0 + l];
}

template <int MODEL, int NEQ, int NCONS, int OCC>
int32_t launch_stream_t(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    auto kern = rhs_stream_kernel<MODEL, NEQ, NCONS, OCC>;
    StreamArgs a;
    a.packs = h->d_packs;
    a.dir = h->d_pack_dir;
    a.list = list;
    a.list_off = off;
    a.count = count;
    a.pack_cap = h->pack_cap;
    a.u_cap = (h->max_nloc * NEQ + 2 + 1) & ~1;
    a.plane = 3 * h->dm.tile_tris + 16;
    static const bool pf = getenv("FVM_STREAM_PREFETCH") != nullptr;  // measured at 4096^2: no gain (0.349 vs 0.341 ms), off
    a.prefetch = pf ? 1 : 0;
    // Two contribution buffers need one barrier per tile; ONE buffer needs a second barrier but halves the planes.  The
    // second form is taken only where it buys a resident CTA (systems: 2 -> 3 CTAs/SM, 2-species kernel 0.750 -> 0.703 ms;
    // scalar kernels sit at 3 either way and lose 3 % to the extra barrier, gpurun_out/r3c_planes_ab.log).
    const int32_t stage_smem = (int32_t)(SK_STAGES * (a.pack_cap + a.u_cap * sizeof(double)) + 2 * SK_STAGES * sizeof(uint64_t));
    const int32_t plane_smem = (int32_t)(NEQ * a.plane * sizeof(double));
    int32_t& configured = h->smem_configured[(const void*)kern];  // dynamic shared memory the choice below was made for
    int32_t& occ = h->occ_cache[(const void*)kern];
    if (2 * plane_smem + stage_smem <= 227 * 1024 ? configured != 2 * plane_smem + stage_smem && configured != plane_smem + stage_smem
                                                  : configured != plane_smem + stage_smem) {
        if (plane_smem + stage_smem > 227 * 1024)
            return fvm_fail(h, FVM_ERR_ARG, "streaming RHS kernel: tile needs more than 227 KB of shared memory; lower tile_triangles");
        // the opt-in is per function and device, not per handle: always the hardware maximum
        FVM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int nb1 = 0, nb2 = 0;
        FVM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, kern, NCONS + SK_PROD, plane_smem + stage_smem));
        if (2 * plane_smem + stage_smem <= 227 * 1024)
            FVM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, kern, NCONS + SK_PROD, 2 * plane_smem + stage_smem));
        if (nb1 < 1) return fvm_fail(h, FVM_ERR_CUDA, "streaming RHS kernel does not fit on an SM");
        static const char* nb_env = getenv("FVM_STREAM_PLANES");  // experiment knob: force 1 or 2 buffers
        const bool one = nb_env ? nb_env[0] == '1' : nb1 > nb2;
        configured = (one || nb2 < 1 ? 1 : 2) * plane_smem + stage_smem;
        occ = (one || nb2 < 1) ? nb1 : nb2;
    }
    a.nbuf = configured == plane_smem + stage_smem ? 1 : 2;
    if (getenv("FVM_STREAM_VERBOSE")) {
        static int said = 0;
        if (said++ < 4)
            fprintf(stderr, "[fvm_stream] model %d neq %d: %d consumer threads, %d-triangle tiles, %d contribution buffer(s), %d B shared memory, %d CTAs/SM\n",
                    MODEL, NEQ, NCONS, h->dm.tile_tris, a.nbuf, configured, occ);
    }
    const int32_t smem = configured;
    h->smem_rhs = smem;
    const int grid = std::min(count, h->sm_count * occ);
    cudaStream_t st = h->launch_stream;
    if (st == h->stream) fvm_prof_begin(h);
    kern<<<grid, NCONS + SK_PROD, smem, st>>>(h->dm, h->flux, h->source, t, u, du, a);
    if (st == h->stream) fvm_prof_end(h);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

template <int MODEL, int NEQ>
int32_t launch_stream_threads(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    // Thread / tile balance (gpurun_out/s3f..s3i_sweep*.log, 4096^2): with T triangles per tile and W consumer warps the
    // triangle pass takes ceil(T / 32W) iterations and the node pass ceil(units / W) rounds of ~T/64 + 2 units.  The round-2
    // default (T = 512, W = 8) fills 10 of 16 node-pass slots; T = 768 = 3 x 256 fills three whole triangle iterations and 14 of
    // 16 slots and, with ONE contribution buffer, still fits 3 CTAs per SM: const-D 0.371 -> 0.347 ms per RHS (kernel 0.330 ->
    // 0.310), u-dependent flux 0.548 -> 0.536.  Smaller CTAs with the same balance come close (T = 576, W = 6, 4 CTAs: 0.355 /
    // 0.311) but leave more interface nodes; W = 7 (T = 512, 4 CTAs, 60 registers) 0.362; systems stay at T = 512 (0.756 vs 0.755).
    int threads = h->stream_threads;
    if (threads == 0) threads = 256;
    switch (threads) {
        case 512: return launch_stream_t<MODEL, NEQ, 512, 2>(h, t, u, du, list, off, count);
        case 384: return launch_stream_t<MODEL, NEQ, 384, 2>(h, t, u, du, list, off, count);
        case 224:  // 7 consumer warps + the producer = 256 threads: 4 CTAs/SM at 60-64 registers
            if constexpr (NEQ == 1) return launch_stream_t<MODEL, NEQ, 224, 4>(h, t, u, du, list, off, count);
            return launch_stream_t<MODEL, NEQ, 256, 2>(h, t, u, du, list, off, count);
        case 192:  // 6 consumer warps: 4 CTAs/SM with the full register budget (systems: 3)
            return launch_stream_t<MODEL, NEQ, 192, (NEQ >= 2 ? 3 : 4)>(h, t, u, du, list, off, count);
        default:
            if (h->stream_occ == 4) return launch_stream_t<MODEL, NEQ, 256, 4>(h, t, u, du, list, off, count);
            // systems: a register budget for 2 CTAs/SM (90 registers, no spills, two contribution buffers) beats 3 CTAs/SM with
            // 72 registers, spills and one buffer: 2-species Keller-Segel 0.703 -> 0.685 ms (gpurun_out/r3j_occ2_ab.log);
            // scalar kernels are the other way round (u-dependent flux 0.500 vs 0.511 ms)
            if (h->stream_occ == 2 || (h->stream_occ == 0 && NEQ >= 2)) return launch_stream_t<MODEL, NEQ, 256, 2>(h, t, u, du, list, off, count);
            return launch_stream_t<MODEL, NEQ, 256, 3>(h, t, u, du, list, off, count);
    }
}

template <int NEQ>
int32_t launch_stream_neq(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    switch (h->flux.model) {
        case FVM_FLUX_DIFF_CONST: return launch_stream_threads<FVM_FLUX_DIFF_CONST, NEQ>(h, t, u, du, list, off, count);
        case FVM_FLUX_DIFF_TABLE: return launch_stream_threads<FVM_FLUX_DIFF_TABLE, NEQ>(h, t, u, du, list, off, count);
        case FVM_FLUX_DIFF_POWER: return launch_stream_threads<FVM_FLUX_DIFF_POWER, NEQ>(h, t, u, du, list, off, count);
        case FVM_FLUX_ADVDIFF: return launch_stream_threads<FVM_FLUX_ADVDIFF, NEQ>(h, t, u, du, list, off, count);
        case FVM_FLUX_KELLER_SEGEL:
            if constexpr (NEQ == 2) return launch_stream_threads<FVM_FLUX_KELLER_SEGEL, 2>(h, t, u, du, list, off, count);
            return fvm_fail(h, FVM_ERR_ARG, "Keller-Segel flux needs neq == 2");
    }
    return fvm_fail(h, FVM_ERR_UNSUPPORTED, "flux model is not in the compiled registry");
}

}  // namespace

int32_t fvm_launch_rhs_stream(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    if (count <= 0) return FVM_OK;
    switch (h->neq) {
        case 1: return launch_stream_neq<1>(h, t, u, du, list, off, count);
        case 2: return launch_stream_neq<2>(h, t, u, du, list, off, count);
        case 3: return launch_stream_neq<3>(h, t, u, du, list, off, count);
        case 4: return launch_stream_neq<4>(h, t, u, du, list, off, count);
    }
    return fvm_fail(h, FVM_ERR_ARG, "unsupported neq");
}

int32_t fvm_stream_fill_vinv(fvm_ctx* h) {
    pack_vinv_kernel<<<h->dm.n_tiles, 128, 0, h->stream>>>(h->d_packs, h->d_pack_dir, h->dm.vol);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}
