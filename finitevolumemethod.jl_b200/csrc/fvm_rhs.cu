// libfvmcuda: the semi-discrete right-hand side  du = fvm_eqs!(du, u, p, t)
// (/root/reference/src/equations/main_equations.jl:38-44) as three kernels:
//
//   rhs_tile_kernel      one CTA per tile of TT Hilbert-sorted triangles: stages the tile's node
//                        values in shared memory, streams the triangle geometry SoA once with
//                        coalesced loads, evaluates the three control-volume-edge fluxes per
//                        triangle (triangle_contributions.jl:28-35), then gathers them per node in
//                        a fixed order from shared memory.  Interior nodes are finished in place
//                        (node pass, source_contributions.jl:33-68) and `du` is written once;
//                        interface nodes emit one partial sum per tile.
//   rhs_boundary_kernel  live boundary edges (boundary_edge_contributions.jl:2-86) -> partials.
//   rhs_interface_kernel sums the partials of interface/boundary nodes in slot order, node pass.
//
// No atomics anywhere: every output word has exactly one writer and a fixed summation order.
#include <cuda_pipeline.h>

#include <chrono>

#include "fvm_device.cuh"

#define RHS_BLOCK 256
#define MODEL_VOLUME 100

// ------------------------------------------------------------------------------------------
// register budget (CTAs of 256 threads per SM): the kernels that stream the full SoA keep 3 resident
// (<= 85 registers; measured best), the reduced-stream kernel wants 8 small CTAs, recompute kernels 4
template <int MODEL, int NEQ>
struct TileOcc {
    static constexpr bool vol = (MODEL == MODEL_VOLUME);
    static constexpr bool full = FluxTraits<vol ? 0 : MODEL>::full;
    template <int GEOM>
    static constexpr int min_blocks() {
        if (vol) return 1;
        if (NEQ == 1) return full ? 3 : (GEOM == 1 ? 4 : 8);
        if (NEQ == 2) return 3;
        return 2;
    }
};
template <int MODEL, int NEQ, int GEOM, int OCC = 0>
__global__ void __launch_bounds__(RHS_BLOCK, OCC ? OCC : TileOcc<MODEL, NEQ>::template min_blocks<GEOM>())
    rhs_tile_kernel(const DevMesh m, const FluxParams fp, const SourceParams sp, const double t,
                    const double* __restrict__ u, double* __restrict__ du, const int smem_nloc,
                    const int32_t* __restrict__ tile_list, const int tile_off) {
    extern __shared__ double smem[];
    constexpr bool VOL = (MODEL == MODEL_VOLUME);
    constexpr bool NEED_XY = VOL || GEOM == 1;
    const int TT = m.tile_tris;
    double* u_s = smem;                                            // [smem_nloc][NEQ]
    double* xy_s = u_s + (VOL ? 0 : (size_t)smem_nloc * NEQ);      // [smem_nloc][2]
    double* c_s = xy_s + (NEED_XY ? 2 * (size_t)smem_nloc : 0);    // [3][TT][NEQ]
    uint16_t* inc_s = reinterpret_cast<uint16_t*>(c_s + (size_t)3 * TT * (VOL ? 1 : NEQ));  // [3*TT] gather list
    uint16_t* iptr_s = inc_s + 3 * TT;                                                      // [smem_nloc + 1]

    const int tile = tile_list ? tile_list[blockIdx.x + tile_off] : (int)blockIdx.x;
    const int tid = threadIdx.x;
    const int4 m0 = __ldg(m.tile_meta + 2 * tile), m1 = __ldg(m.tile_meta + 2 * tile + 1);
    const int node0 = m0.x, nint = m0.y, nown = m0.z, nloc = m0.w;
    const int ext0 = m1.x, loc0 = m1.y, pp0 = m1.z, ntri = m1.w;
    const int64_t t0 = (int64_t)tile * TT;

    // ---- stage the tile's gather list early: it is consumed after the second barrier, so its
    // latency overlaps the triangle pass instead of stalling the node pass ----------------------
    {   // cp.async (LDGSTS): no register staging, completion awaited just before the node pass
        const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(m.inc + (size_t)3 * TT * tile);
        uint32_t* dst = reinterpret_cast<uint32_t*>(inc_s);
        for (int k = tid; k < (3 * ntri + 1) / 2; k += RHS_BLOCK) __pipeline_memcpy_async(dst + k, src + k, 4);
        const uint32_t* __restrict__ ip = reinterpret_cast<const uint32_t*>(m.inc_ptr + loc0);
        uint32_t* dp = reinterpret_cast<uint32_t*>(iptr_s);
        for (int k = tid; k < (nloc + 2) / 2; k += RHS_BLOCK) __pipeline_memcpy_async(dp + k, ip + k, 4);
        __pipeline_commit();
    }
    // ---- register prefetch of everything else this thread will need from global memory, so the
    // DRAM latency is paid once per tile (overlapped with the staging) instead of once per use ----
    // (only where registers are cheap: the recompute kernels of flux models without (x,y,u) terms;
    // measured: -16 % there, but the extra registers cost the stored-geometry kernels occupancy)
    constexpr bool PREFETCH = !VOL && GEOM == 1 && !FluxTraits<VOL ? 0 : MODEL>::full;
    constexpr int TPF = 4, NPF = 3;
    ushort4 vpf[TPF];
    double Vpf[NPF];
    uint8_t kpf[NPF][NEQ];
    if constexpr (PREFETCH) {
#pragma unroll
        for (int r = 0; r < TPF; ++r)
            vpf[r] = (tid + r * RHS_BLOCK < ntri) ? __ldg(m.tri_loc + t0 + tid + r * RHS_BLOCK) : make_ushort4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < NPF; ++r) {
            const int l = tid + r * RHS_BLOCK;
            Vpf[r] = l < nint ? __ldg(m.vol + node0 + l) : 1.0;
#pragma unroll
            for (int v = 0; v < NEQ; ++v) kpf[r][v] = l < nint ? __ldg(m.kind + (size_t)v * m.n_nodes + node0 + l) : (uint8_t)0;
        }
    }

    // ---- recompute kernels are bound by the DRAM latency of this staging phase (ncu: 45 % of the warp samples
    // wait here on a long scoreboard), so every CTA pulls the lists and node ranges of the tile that will be
    // scheduled about one CTA lifetime later into L2: its loads then hit L2 instead of DRAM -------------------
    if constexpr (PREFETCH) {
        const int nt = tile + m.pf_ahead;
        if (m.pf_ahead > 0 && !tile_list && nt < m.n_tiles) {
            const int4 p0 = __ldg(m.tile_meta + 2 * nt), p1 = __ldg(m.tile_meta + 2 * nt + 1);
            auto pf = [&](const void* base, int bytes) {
                const char* b = reinterpret_cast<const char*>(base);
                for (int off = tid * 128; off < bytes; off += RHS_BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + off));
            };
            pf(u + (size_t)p0.x * NEQ, p0.z * NEQ * 8);
            pf(m.xy + (size_t)p0.x * 2, p0.z * 16);
            pf(m.tri_loc + (int64_t)nt * TT, p1.w * 8);
            pf(m.inc + (size_t)3 * TT * nt, p1.w * 6);
            pf(m.inc_ptr + p1.y, (p0.w + 1) * 2);
            pf(m.ext_ids + p1.x, (p0.w - p0.z) * 4);
            pf(m.vol + p0.x, p0.y * 8);
        }
    }
    // ---- stage node data: own range is contiguous, external interface nodes are gathered ----
    if constexpr (!VOL) {
        for (int idx = tid; idx < nown * NEQ; idx += RHS_BLOCK) u_s[idx] = u[(size_t)node0 * NEQ + idx];
        for (int k = tid; k < nloc - nown; k += RHS_BLOCK) {
            const int g = m.ext_ids[ext0 + k];
#pragma unroll
            for (int v = 0; v < NEQ; ++v) u_s[(nown + k) * NEQ + v] = u[(size_t)g * NEQ + v];
        }
    }
    if constexpr (NEED_XY) {
        for (int idx = tid; idx < nown * 2; idx += RHS_BLOCK) xy_s[idx] = m.xy[(size_t)node0 * 2 + idx];
        for (int k = tid; k < nloc - nown; k += RHS_BLOCK) {
            const int g = m.ext_ids[ext0 + k];
            xy_s[(nown + k) * 2] = m.xy[2 * (size_t)g];
            xy_s[(nown + k) * 2 + 1] = m.xy[2 * (size_t)g + 1];
        }
    }
    __syncthreads();

    // ---- triangle pass ------------------------------------------------------------------------
    const int n_iter = (ntri - tid + RHS_BLOCK - 1) / RHS_BLOCK;
#pragma unroll 1
    for (int r = 0; r < n_iter; ++r) {
        const int lt = tid + r * RHS_BLOCK;
        const int64_t gt = t0 + lt;
        ushort4 vv;
        if constexpr (PREFETCH) {
            switch (r) {  // prefetched ids for the first TPF rounds
                case 0: vv = vpf[0]; break;
                case 1: vv = vpf[1]; break;
                case 2: vv = vpf[2]; break;
                case 3: vv = vpf[3]; break;
                default: vv = m.tri_loc[gt]; break;
            }
        } else {
            vv = m.tri_loc[gt];
        }
        if constexpr (VOL) {
            TriGeom G;
            double S[3];
            tri_geometry<true>(xy_s[2 * vv.x], xy_s[2 * vv.x + 1], xy_s[2 * vv.y], xy_s[2 * vv.y + 1], xy_s[2 * vv.z],
                               xy_s[2 * vv.z + 1], G, S);
            c_s[0 * TT + lt] = S[0];
            c_s[1 * TT + lt] = S[1];
            c_s[2 * TT + lt] = S[2];
        } else {
            constexpr bool FULL = FluxTraits<MODEL>::full;
            double s[9], mx[3], my[3], nlx[3], nly[3], dt3[3] = {0, 0, 0};
            if constexpr (GEOM == 0) {
                const double* __restrict__ geo = m.geo + gt;
                const int64_t st = m.tpad;
#pragma unroll
                for (int c = 0; c < (FULL ? 9 : 6); ++c) s[c] = __ldg(geo + c * st);
                if constexpr (FULL) {
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        mx[e] = __ldg(geo + (9 + 2 * e) * st);
                        my[e] = __ldg(geo + (10 + 2 * e) * st);
                    }
                }
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    nlx[e] = __ldg(geo + (15 + 2 * e) * st);
                    nly[e] = __ldg(geo + (16 + 2 * e) * st);
                }
            } else {
                TriGeom G;
                tri_geometry<false, FULL>(xy_s[2 * vv.x], xy_s[2 * vv.x + 1], xy_s[2 * vv.y], xy_s[2 * vv.y + 1],
                                          xy_s[2 * vv.z], xy_s[2 * vv.z + 1], G, nullptr);
#pragma unroll
                for (int c = 0; c < 9; ++c) s[c] = G.s[c];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    mx[e] = G.mx[e];
                    my[e] = G.my[e];
                    nlx[e] = G.ey[e];
                    nly[e] = -G.ex[e];
                }
            }
            if constexpr (FluxTraits<MODEL>::table) {
#pragma unroll
                for (int e = 0; e < 3; ++e) dt3[e] = __ldg(m.dtab + e * m.tpad + gt);
            }
            // shape-function coefficients, shape_functions.jl:2-19
            double a[NEQ], b[NEQ], g[NEQ];
#pragma unroll
            for (int v = 0; v < NEQ; ++v) {
                const double ui = u_s[vv.x * NEQ + v], uj = u_s[vv.y * NEQ + v], uk = u_s[vv.z * NEQ + v];
                if constexpr (FULL) {  // fluxes that read u = alpha x + beta y + gamma: the reference's roundings throughout
                    a[v] = shape_coeff(s[0], s[1], s[2], ui, uj, uk);
                    b[v] = shape_coeff(s[3], s[4], s[5], ui, uj, uk);
                    g[v] = shape_coeff(s[6], s[7], s[8], ui, uj, uk);
                } else {
                    a[v] = s[0] * ui + s[1] * uj + s[2] * uk;
                    b[v] = s[3] * ui + s[4] * uj + s[5] * uk;
                    g[v] = 0.0;
                }
            }
            double Q[3][NEQ];
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double qx[NEQ], qy[NEQ];
                flux_eval<MODEL, NEQ>(fp, FULL ? mx[e] : 0.0, FULL ? my[e] : 0.0, t, a, b, g, dt3[e], qx, qy);
#pragma unroll
                for (int v = 0; v < NEQ; ++v) Q[e][v] = qx[v] * nlx[e] + qy[v] * nly[e];
            }
            // vertex contributions, triangle_contributions.jl:10-25
#pragma unroll
            for (int v = 0; v < NEQ; ++v) {
                c_s[(0 * TT + lt) * NEQ + v] = Q[2][v] - Q[0][v];
                c_s[(1 * TT + lt) * NEQ + v] = Q[0][v] - Q[1][v];
                c_s[(2 * TT + lt) * NEQ + v] = Q[1][v] - Q[2][v];
            }
        }
    }
    __pipeline_wait_prior(0);
    __syncthreads();

    // ---- per-node gather in ascending triangle order, then node pass --------------------------
    const uint16_t* iptr = iptr_s;
    const uint16_t* inc = inc_s;
    int round = 0;
    for (int l = tid; l < nloc; l += RHS_BLOCK, ++round) {
        const int beg = iptr[l], end = iptr[l + 1];
        double acc[NEQ];
#pragma unroll
        for (int v = 0; v < NEQ; ++v) acc[v] = 0.0;
        for (int e = beg; e < end; ++e) {
            const int code = inc[e];
            const int off = ((code & 3) * TT + (code >> 2)) * NEQ;
#pragma unroll
            for (int v = 0; v < NEQ; ++v) acc[v] += c_s[off + v];
        }
        if (l < nint) {
            if constexpr (VOL) du[node0 + l] = acc[0];
            else if constexpr (!PREFETCH) node_finish<NEQ>(m, sp, t, node0 + l, acc, u_s + l * NEQ, du);
            else {
                switch (round) {
                    case 0: node_finish<NEQ>(m, sp, t, node0 + l, acc, u_s + l * NEQ, du, Vpf[0], kpf[0]); break;
                    case 1: node_finish<NEQ>(m, sp, t, node0 + l, acc, u_s + l * NEQ, du, Vpf[1], kpf[1]); break;
                    case 2: node_finish<NEQ>(m, sp, t, node0 + l, acc, u_s + l * NEQ, du, Vpf[2], kpf[2]); break;
                    default: node_finish<NEQ>(m, sp, t, node0 + l, acc, u_s + l * NEQ, du); break;
                }
            }
        } else {
            const size_t p = (size_t)m.ppos[pp0 + (l - nint)] * NEQ;
#pragma unroll
            for (int v = 0; v < NEQ; ++v) m.partial[p + v] = acc[v];
        }
    }
}

// ------------------------------------------------------------------------------------------
template <int NEQ, bool VOL>
__global__ void __launch_bounds__(256)
    rhs_interface_kernel(const DevMesh m, const SourceParams sp, const double t, const double* __restrict__ u,
                         double* __restrict__ du, const int32_t* __restrict__ list = nullptr, const int list_off = 0,
                         const int list_count = 0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (list) {  // explicit subset of the interface nodes (host-buffer pipeline)
        if (k >= list_count) return;
        k = list[list_off + k];
    }
    if (k < m.n_ifc) {
        const int g = m.ifc_node[k];
        const int beg = m.ifc_pptr[k], end = m.ifc_pptr[k + 1];
        double acc[NEQ];
#pragma unroll
        for (int v = 0; v < NEQ; ++v) acc[v] = 0.0;
        for (int p = beg; p < end; ++p) {
#pragma unroll
            for (int v = 0; v < NEQ; ++v) acc[v] += m.partial[(size_t)p * NEQ + v];
        }
        if constexpr (VOL) {
            du[g] = acc[0];
        } else {
            double uv[NEQ];
#pragma unroll
            for (int v = 0; v < NEQ; ++v) uv[v] = u[(size_t)g * NEQ + v];
            node_finish<NEQ>(m, sp, t, g, acc, uv, du);
        }
    } else {
        // points that are not vertices of any triangle: du = 0 (source_contributions.jl:36-37)
        const int g = m.n_vertices + (k - m.n_ifc);
        if (g < m.n_nodes) {
            if constexpr (VOL) du[g] = 0.0;
            else {
#pragma unroll
                for (int v = 0; v < NEQ; ++v) du[(size_t)g * NEQ + v] = 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
template <int NEQ>
__device__ __forceinline__ void flux_dispatch(const FluxParams& fp, double x, double y, double t, const double* a,
                                              const double* b, const double* g, double dtab, double* qx, double* qy) {
    switch (fp.model) {
        case FVM_FLUX_DIFF_TABLE: flux_eval<FVM_FLUX_DIFF_TABLE, NEQ>(fp, x, y, t, a, b, g, dtab, qx, qy); break;
        case FVM_FLUX_DIFF_POWER: flux_eval<FVM_FLUX_DIFF_POWER, NEQ>(fp, x, y, t, a, b, g, dtab, qx, qy); break;
        case FVM_FLUX_ADVDIFF: flux_eval<FVM_FLUX_ADVDIFF, NEQ>(fp, x, y, t, a, b, g, dtab, qx, qy); break;
        case FVM_FLUX_KELLER_SEGEL:
            if constexpr (NEQ == 2) flux_eval<FVM_FLUX_KELLER_SEGEL, NEQ>(fp, x, y, t, a, b, g, dtab, qx, qy);
            break;
        default: flux_eval<FVM_FLUX_DIFF_CONST, NEQ>(fp, x, y, t, a, b, g, dtab, qx, qy); break;
    }
}

// boundary_edge_contributions.jl:41-86 and control_volumes.jl:41-56
template <int NEQ>
__global__ void __launch_bounds__(128)
    rhs_boundary_kernel(const DevMesh m, const FluxParams fp, const double t, const BndEdge* __restrict__ edges,
                        const double* __restrict__ dbnd, const int n_edges, const double* __restrict__ u,
                        const int32_t* __restrict__ list = nullptr, const int list_off = 0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edges) return;
    if (list) k = list[list_off + k];  // explicit subset (host-buffer pipeline); n_edges is then the subset size
    const BndEdge E = edges[k];
    TriGeom G;
    tri_geometry<true>(m.xy[2 * (size_t)E.v[0]], m.xy[2 * (size_t)E.v[0] + 1], m.xy[2 * (size_t)E.v[1]],
                       m.xy[2 * (size_t)E.v[1] + 1], m.xy[2 * (size_t)E.v[2]], m.xy[2 * (size_t)E.v[2] + 1], G, nullptr);
    double a[NEQ], b[NEQ], g[NEQ];
#pragma unroll
    for (int v = 0; v < NEQ; ++v) {
        const double ui = u[(size_t)E.v[0] * NEQ + v], uj = u[(size_t)E.v[1] * NEQ + v], uk = u[(size_t)E.v[2] * NEQ + v];
        a[v] = shape_coeff(G.s[0], G.s[1], G.s[2], ui, uj, uk);
        b[v] = shape_coeff(G.s[3], G.s[4], G.s[5], ui, uj, uk);
        g[v] = shape_coeff(G.s[6], G.s[7], G.s[8], ui, uj, uk);
    }
    const double dx = E.qx - E.px, dy = E.qy - E.py;
    const double lij = sqrt(dx * dx + dy * dy);
    const double nx = dy / lij, ny = -dx / lij;
    const double mijx = (E.px + E.qx) / 2, mijy = (E.py + E.qy) / 2;
    const double ptx[2] = {(E.px + mijx) / 2, (E.qx + mijx) / 2};
    const double pty[2] = {(E.py + mijy) / 2, (E.qy + mijy) / 2};
    const double hx = mijx - E.px, hy = mijy - E.py;
    const double l = sqrt(hx * hx + hy * hy);
    const int slot[2] = {E.slot_i, E.slot_j};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        double qx[NEQ], qy[NEQ];
        flux_dispatch<NEQ>(fp, ptx[q], pty[q], t, a, b, g, dbnd ? dbnd[2 * k + q] : 0.0, qx, qy);
#pragma unroll
        for (int v = 0; v < NEQ; ++v) {
            double Q;
            if (E.kind[v] == FVM_EDGE_NEUMANN) {
                const CondFn c = m.cond[v * FVM_MAX_COND_FN + E.fidx[v]];
                const double ushape = shape_value(a[v], b[v], g[v], ptx[q], pty[q]);
                Q = cond_eval(c, ptx[q], pty[q], t, ushape) * l;
            } else {
                Q = (qx[v] * nx + qy[v] * ny) * l;
            }
            m.partial[(size_t)slot[q] * NEQ + v] = -Q;  // du[i] -= Q_i
        }
    }
}

// update_dirichlet_nodes!, dirichlet.jl:2-86
__global__ void dirichlet_kernel(const DevMesh m, const int32_t* __restrict__ pairs, const int n_pairs, const double t,
                                 double* __restrict__ u) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const int g = pairs[2 * k], v = pairs[2 * k + 1];
    const CondFn c = m.cond[v * FVM_MAX_COND_FN + m.fidx[(size_t)v * m.n_nodes + g]];
    const size_t idx = (size_t)g * m.neq + v;
    u[idx] = cond_eval(c, m.xy[2 * (size_t)g], m.xy[2 * (size_t)g + 1], t, u[idx]);
}

// caller order <-> native order
__global__ void permute_kernel(const int32_t* __restrict__ old_of_new, const double* __restrict__ src,
                               double* __restrict__ dst, const int64_t n, const int neq, const int to_native) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * neq) return;
    const int64_t g = idx / neq;
    const int v = (int)(idx - g * neq);
    const int64_t o = (int64_t)old_of_new[g] * neq + v;
    if (to_native) dst[idx] = src[o];
    else dst[o] = src[idx];
}

// stored geometry SoA: one thread per native triangle, exact (reference-order) arithmetic
__global__ void geometry_kernel(const DevMesh m, const int32_t* __restrict__ tri_native, double* __restrict__ geo) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= m.n_tris) return;
    const int i = tri_native[3 * gt], j = tri_native[3 * gt + 1], k = tri_native[3 * gt + 2];
    TriGeom G;
    tri_geometry<true>(m.xy[2 * (size_t)i], m.xy[2 * (size_t)i + 1], m.xy[2 * (size_t)j], m.xy[2 * (size_t)j + 1],
                       m.xy[2 * (size_t)k], m.xy[2 * (size_t)k + 1], G, nullptr);
    const int64_t st = m.tpad;
#pragma unroll
    for (int c = 0; c < 9; ++c) geo[c * st + gt] = G.s[c];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        geo[(9 + 2 * e) * st + gt] = G.mx[e];
        geo[(10 + 2 * e) * st + gt] = G.my[e];
        geo[(15 + 2 * e) * st + gt] = G.ey[e];
        geo[(16 + 2 * e) * st + gt] = -G.ex[e];
    }
}

// read-back of TriangleProperties (geometry.jl:21-26) in the caller's triangle order
__global__ void export_geometry_kernel(const DevMesh m, const int32_t* __restrict__ tri_native,
                                       const int32_t* __restrict__ tri_old_of_new, double* s9, double* mid6, double* nrm6,
                                       double* len3) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= m.n_tris) return;
    const int i = tri_native[3 * gt], j = tri_native[3 * gt + 1], k = tri_native[3 * gt + 2];
    TriGeom G;
    tri_geometry<true>(m.xy[2 * (size_t)i], m.xy[2 * (size_t)i + 1], m.xy[2 * (size_t)j], m.xy[2 * (size_t)j + 1],
                       m.xy[2 * (size_t)k], m.xy[2 * (size_t)k + 1], G, nullptr);
    const int64_t o = tri_old_of_new[gt];
    if (s9)
        for (int c = 0; c < 9; ++c) s9[9 * o + c] = G.s[c];
    for (int e = 0; e < 3; ++e) {
        if (mid6) {
            mid6[6 * o + 2 * e] = G.mx[e];
            mid6[6 * o + 2 * e + 1] = G.my[e];
        }
        // geometry.jl:153-161: l = norm(e); n = (e_y / l, -e_x / l)
        const double l = __dsqrt_rn(__dadd_rn(__dmul_rn(G.ex[e], G.ex[e]), __dmul_rn(G.ey[e], G.ey[e])));
        if (len3) len3[3 * o + e] = l;
        if (nrm6) {
            nrm6[6 * o + 2 * e] = __ddiv_rn(G.ey[e], l);
            nrm6[6 * o + 2 * e + 1] = __ddiv_rn(-G.ex[e], l);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int32_t rhs_smem_bytes(const fvm_ctx* h, int neq, bool vol, bool need_xy) {
    size_t d = 0;
    if (!vol) d += (size_t)h->max_nloc * neq;
    if (need_xy) d += 2 * (size_t)h->max_nloc;
    d += (size_t)3 * h->dm.tile_tris * (vol ? 1 : neq);
    size_t bytes = d * sizeof(double);
    bytes += sizeof(uint16_t) * ((size_t)3 * h->dm.tile_tris + h->max_nloc + 4);  // staged gather list
    return (int32_t)((bytes + 15) & ~(size_t)15);
}

template <int MODEL, int NEQ, int GEOM>
static int32_t launch_tile(fvm_ctx* h, double t, const double* u, double* du, int part = 0) {
    constexpr bool VOL = (MODEL == MODEL_VOLUME);
    const int32_t smem = rhs_smem_bytes(h, NEQ, VOL, VOL || GEOM == 1);
    auto kern = rhs_tile_kernel<MODEL, NEQ, GEOM>;
    if constexpr (NEQ == 2 && MODEL == FVM_FLUX_KELLER_SEGEL && GEOM == 0) {
        static const char* e = getenv("FVM_SYS_MINB");  // experiment knob: 4 resident CTAs (<= 64 registers)
        if (e && e[0] == '4') kern = rhs_tile_kernel<MODEL, NEQ, GEOM, 4>;
    }
    if (smem > 200 * 1024) return fvm_fail(h, FVM_ERR_ARG, "tile needs more than 200 KB of shared memory; lower tile_triangles");
    int32_t& configured = h->smem_configured[(const void*)kern];
    if (configured < smem) {
        FVM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    h->smem_rhs = smem;
    int grid = h->dm.n_tiles, off = 0;
    const int32_t* list = nullptr;
    if (part == 1) {
        list = h->d_tile_order;
        grid = h->n_tiles_indep;
    } else if (part == 2) {
        list = h->d_tile_order;
        off = h->n_tiles_indep;
        grid = h->dm.n_tiles - h->n_tiles_indep;
    } else if (part == 4) {
        list = h->pipe_list;
        off = h->pipe_off;
        grid = h->pipe_count;
    }
    if (grid == 0) return FVM_OK;
    cudaStream_t st = h->launch_stream;
    if (st == h->stream) fvm_prof_begin(h);
    kern<<<grid, RHS_BLOCK, smem, st>>>(h->dm, h->flux, h->source, t, u, du, h->max_nloc, list, off);
    if (st == h->stream) fvm_prof_end(h);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

template <int NEQ>
static int32_t launch_rhs_neq(fvm_ctx* h, double t, const double* u, double* du, int part) {
    int32_t rc = FVM_OK;
    if (h->n_bnd_live > 0 && (part == 0 || part == 2)) {
        rhs_boundary_kernel<NEQ><<<(h->n_bnd_live + 127) / 128, 128, 0, h->launch_stream>>>(h->dm, h->flux, t, h->d_bnd, h->d_dbnd,
                                                                                            h->n_bnd_live, u);
        FVM_CUDA(h, cudaGetLastError());
    }
    if (part == 3) {
        const int n_tail3 = h->dm.n_ifc + (h->dm.n_nodes - h->dm.n_vertices);
        if (n_tail3 > 0) {
            rhs_interface_kernel<NEQ, false><<<(n_tail3 + 255) / 256, 256, 0, h->launch_stream>>>(h->dm, h->source, t, u, du);
            FVM_CUDA(h, cudaGetLastError());
        }
        return FVM_OK;
    }
    const int model = h->flux.model;
    const int geom = h->geometry_mode;
    if (geom == 1 && h->packs_ready) {  // persistent streaming kernel (fvm_rhs_stream.cu)
        int count = h->dm.n_tiles, off = 0;
        const int32_t* list = nullptr;
        if (part == 1) {
            list = h->d_tile_order;
            count = h->n_tiles_indep;
        } else if (part == 2) {
            list = h->d_tile_order;
            off = h->n_tiles_indep;
            count = h->dm.n_tiles - h->n_tiles_indep;
        } else if (part == 4) {
            list = h->pipe_list;
            off = h->pipe_off;
            count = h->pipe_count;
        }
        if ((rc = fvm_launch_rhs_stream(h, t, u, du, list, off, count))) return rc;
        const int n_tail_s = h->dm.n_ifc + (h->dm.n_nodes - h->dm.n_vertices);
        if (n_tail_s > 0 && part == 0) {
            rhs_interface_kernel<NEQ, false><<<(n_tail_s + 255) / 256, 256, 0, h->launch_stream>>>(h->dm, h->source, t, u, du);
            FVM_CUDA(h, cudaGetLastError());
        }
        return FVM_OK;
    }
#define TILE_CASE(M)                                                        \
    case M:                                                                 \
        rc = geom ? launch_tile<M, NEQ, 1>(h, t, u, du, part) : launch_tile<M, NEQ, 0>(h, t, u, du, part); \
        break;
    switch (model) {
        TILE_CASE(FVM_FLUX_DIFF_CONST)
        TILE_CASE(FVM_FLUX_DIFF_TABLE)
        TILE_CASE(FVM_FLUX_DIFF_POWER)
        TILE_CASE(FVM_FLUX_ADVDIFF)
        case FVM_FLUX_KELLER_SEGEL:
            if constexpr (NEQ == 2) {
                rc = geom ? launch_tile<FVM_FLUX_KELLER_SEGEL, 2, 1>(h, t, u, du, part)
                          : launch_tile<FVM_FLUX_KELLER_SEGEL, 2, 0>(h, t, u, du, part);
            } else {
                rc = fvm_fail(h, FVM_ERR_ARG, "Keller-Segel flux needs neq == 2");
            }
            break;
        default: rc = fvm_fail(h, FVM_ERR_UNSUPPORTED, "flux model is not in the compiled registry");
    }
#undef TILE_CASE
    if (rc) return rc;
    const int n_tail = h->dm.n_ifc + (h->dm.n_nodes - h->dm.n_vertices);
    if (n_tail > 0 && part == 0) {
        rhs_interface_kernel<NEQ, false><<<(n_tail + 255) / 256, 256, 0, h->launch_stream>>>(h->dm, h->source, t, u, du);
        FVM_CUDA(h, cudaGetLastError());
    }
    return FVM_OK;
}

int32_t fvm_launch_rhs_part(fvm_ctx* h, double t, const double* u, double* du, int part) {
    switch (h->neq) {
        case 1: return launch_rhs_neq<1>(h, t, u, du, part);
        case 2: return launch_rhs_neq<2>(h, t, u, du, part);
        case 3: return launch_rhs_neq<3>(h, t, u, du, part);
        case 4: return launch_rhs_neq<4>(h, t, u, du, part);
    }
    return fvm_fail(h, FVM_ERR_ARG, "unsupported neq");
}

int32_t fvm_launch_rhs(fvm_ctx* h, double t, const double* u, double* du) { return fvm_launch_rhs_part(h, t, u, du, 0); }

template <int NEQ>
static int32_t launch_interface_list(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    rhs_interface_kernel<NEQ, false><<<(count + 255) / 256, 256, 0, h->launch_stream>>>(h->dm, h->source, t, u, du, list, off, count);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

int32_t fvm_launch_rhs_interface_list(fvm_ctx* h, double t, const double* u, double* du, const int32_t* list, int off, int count) {
    if (count <= 0) return FVM_OK;
    switch (h->neq) {
        case 1: return launch_interface_list<1>(h, t, u, du, list, off, count);
        case 2: return launch_interface_list<2>(h, t, u, du, list, off, count);
        case 3: return launch_interface_list<3>(h, t, u, du, list, off, count);
        case 4: return launch_interface_list<4>(h, t, u, du, list, off, count);
    }
    return fvm_fail(h, FVM_ERR_ARG, "unsupported neq");
}

template <int NEQ>
static int32_t launch_boundary_list(fvm_ctx* h, double t, const double* u, const int32_t* list, int off, int count) {
    rhs_boundary_kernel<NEQ><<<(count + 127) / 128, 128, 0, h->launch_stream>>>(h->dm, h->flux, t, h->d_bnd, h->d_dbnd, count, u, list, off);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

int32_t fvm_launch_rhs_boundary_list(fvm_ctx* h, double t, const double* u, const int32_t* list, int off, int count) {
    if (count <= 0) return FVM_OK;
    switch (h->neq) {
        case 1: return launch_boundary_list<1>(h, t, u, list, off, count);
        case 2: return launch_boundary_list<2>(h, t, u, list, off, count);
        case 3: return launch_boundary_list<3>(h, t, u, list, off, count);
        case 4: return launch_boundary_list<4>(h, t, u, list, off, count);
    }
    return fvm_fail(h, FVM_ERR_ARG, "unsupported neq");
}

// points that are not vertices of any triangle: du = 0 (source_contributions.jl:36-37)
int32_t fvm_launch_rhs_nonvertex(fvm_ctx* h, double* du) {
    const int64_t n = (int64_t)(h->dm.n_nodes - h->dm.n_vertices) * h->neq;
    if (n > 0) FVM_CUDA(h, cudaMemsetAsync(du + (size_t)h->dm.n_vertices * h->neq, 0, sizeof(double) * n, h->launch_stream));
    return FVM_OK;
}

// du = fvm_eqs!(u) with the ghost refresh; the exchange overlaps the tiles that touch no ghost node
int32_t fvm_apply_rhs(fvm_ctx* h, double t, double* x, double* out) {
    if (!h->halo_ready) return fvm_launch_rhs_part(h, t, x, out, 0);
    int32_t rc;
    if (!h->overlap) {
        if ((rc = fvm_halo_exchange(h, x))) return rc;
        return fvm_launch_rhs_part(h, t, x, out, 0);
    }
    // communication stream: exchange -> unpack -> halo-dependent tiles + boundary edges, concurrently with
    // the independent tiles on the compute stream (they fill its tail wave); then the interface kernel
    if ((rc = fvm_halo_begin(h, x))) return rc;
    h->launch_stream = h->comm_stream;
    rc = fvm_launch_rhs_part(h, t, x, out, 2);
    h->launch_stream = h->stream;
    if (rc) return rc;
    if ((rc = fvm_halo_done(h))) return rc;
    if ((rc = fvm_launch_rhs_part(h, t, x, out, 1))) return rc;
    if ((rc = fvm_halo_wait(h))) return rc;
    return fvm_launch_rhs_part(h, t, x, out, 3);
}

int32_t fvm_launch_geometry(fvm_ctx* h, const int32_t* d_tri_native) {
    if (h->geometry_mode == 1) {
        h->dm.geo = nullptr;
        return FVM_OK;
    }
    double* geo = nullptr;
    int32_t rc = fvm_dev_alloc(h, &geo, (size_t)FVM_NGEO * h->dm.tpad);
    if (rc) return rc;
    FVM_CUDA(h, cudaMemsetAsync(geo, 0, sizeof(double) * FVM_NGEO * h->dm.tpad, h->stream));
    h->dm.geo = geo;
    geometry_kernel<<<(unsigned)((h->dm.n_tris + 255) / 256), 256, 0, h->stream>>>(h->dm, d_tri_native, geo);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

int32_t fvm_launch_volumes(fvm_ctx* h) {
    // cv_volumes (geometry.jl:119-135) through the same tile gather: deterministic summation
    double* vol = const_cast<double*>(h->dm.vol);
    FVM_CUDA(h, cudaMemsetAsync(h->dm.partial, 0, sizeof(double) * h->dm.n_partial * h->neq, h->stream));
    int32_t rc = launch_tile<MODEL_VOLUME, 1, 1>(h, 0.0, nullptr, vol, 0);
    if (rc) return rc;
    const int n_tail = h->dm.n_ifc + (h->dm.n_nodes - h->dm.n_vertices);
    if (n_tail > 0) {
        rhs_interface_kernel<1, true><<<(n_tail + 255) / 256, 256, 0, h->stream>>>(h->dm, h->source, 0.0, nullptr, vol);
        FVM_CUDA(h, cudaGetLastError());
    }
    return FVM_OK;
}

int32_t fvm_launch_dirichlet(fvm_ctx* h, double t, double* u) {
    if (h->n_dir == 0) return FVM_OK;
    dirichlet_kernel<<<(h->n_dir + 255) / 256, 256, 0, h->stream>>>(h->dm, h->d_dir_nodes, h->n_dir, t, u);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

int32_t fvm_launch_permute(fvm_ctx* h, const double* src, double* dst, bool to_native) {
    const int64_t n = h->N * h->neq;
    permute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_node_old_of_new, src, dst, h->N, h->neq,
                                                                        to_native ? 1 : 0);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

int32_t fvm_ensure_state(fvm_ctx* h) {
    if (h->d_u) return FVM_OK;
    const size_t n = (size_t)h->N * h->neq;
    int32_t rc;
    if ((rc = fvm_dev_alloc(h, &h->d_u, n))) return rc;
    if ((rc = fvm_dev_alloc(h, &h->d_du, n))) return rc;
    if ((rc = fvm_dev_alloc(h, &h->d_io, n))) return rc;
    return FVM_OK;
}

#define NEED_FINAL(h)                                                                       \
    do {                                                                                    \
        if (!(h)) return FVM_ERR_ARG;                                                       \
        if (!(h)->finalized) return fvm_fail((h), FVM_ERR_STATE, "call fvm_finalize first"); \
        FVM_CUDA(h, cudaSetDevice((h)->device));                                            \
    } while (0)

// self-check of the recompute path's arithmetic: the contracted / division-free tri_geometry<false> the streaming kernel
// runs must reproduce the individually rounded reference arithmetic (tri_geometry<true>): bit for bit in the variant of
// the u-dependent fluxes; s7..s9, cv-edge midpoints and vectors bit for bit and s1..s6 to one ulp in the other
__global__ void geometry_check_kernel(const DevMesh m, const int32_t* __restrict__ tri_native, unsigned long long* __restrict__ bad) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= m.n_tris) return;
    const int i = tri_native[3 * gt], j = tri_native[3 * gt + 1], k = tri_native[3 * gt + 2];
    TriGeom A, B, C;
    const double px = m.xy[2 * (size_t)i], py = m.xy[2 * (size_t)i + 1], qx = m.xy[2 * (size_t)j], qy = m.xy[2 * (size_t)j + 1];
    const double rx = m.xy[2 * (size_t)k], ry = m.xy[2 * (size_t)k + 1];
    tri_geometry<true>(px, py, qx, qy, rx, ry, A, nullptr);          // the reference's arithmetic
    tri_geometry<false, true>(px, py, qx, qy, rx, ry, B, nullptr);   // u-dependent fluxes
    tri_geometry<false, false>(px, py, qx, qy, rx, ry, C, nullptr);  // fluxes that read alpha and beta only
    bool same = true;
    for (int c = 0; c < 9; ++c) same = same && A.s[c] == B.s[c];
    for (int c = 6; c < 9; ++c) same = same && A.s[c] == C.s[c];
    for (int c = 0; c < 6; ++c) same = same && fabs(A.s[c] - C.s[c]) <= 2.3e-16 * fabs(A.s[c]);  // num * RN(1/D): one ulp
    for (int e = 0; e < 3; ++e)
        same = same && A.mx[e] == B.mx[e] && A.my[e] == B.my[e] && A.ex[e] == B.ex[e] && A.ey[e] == B.ey[e] && A.mx[e] == C.mx[e] &&
               A.my[e] == C.my[e] && A.ex[e] == C.ex[e] && A.ey[e] == C.ey[e];
    if (!same) atomicAdd(bad, 1ULL);
}

extern "C" int32_t fvm_check_recompute_geometry(fvm_handle h, int64_t* n_mismatch) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, n_mismatch, "fvm_check_recompute_geometry: null argument");
    unsigned long long* d = nullptr;
    FVM_CUDA(h, cudaMalloc((void**)&d, sizeof(unsigned long long)));
    cudaMemsetAsync(d, 0, sizeof(unsigned long long), h->stream);
    geometry_check_kernel<<<(unsigned)((h->T + 255) / 256), 256, 0, h->stream>>>(h->dm, h->d_tri_native, d);
    unsigned long long out = 0;
    cudaError_t e = cudaMemcpyAsync(&out, d, sizeof(out), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    FVM_CUDA(h, e);
    *n_mismatch = (int64_t)out;
    return FVM_OK;
}

extern "C" int32_t fvm_rhs_native(fvm_handle h, double t, const double* u, double* du) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u && du, "fvm_rhs_native: null argument");
    return fvm_apply_rhs(h, t, const_cast<double*>(u), du);  // sharded: ghost entries of u are refreshed in place
}

extern "C" int32_t fvm_rhs(fvm_handle h, double t, const double* u, double* du, int32_t on_device) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u && du, "fvm_rhs: null argument");
    int32_t rc = fvm_ensure_state(h);
    if (rc) return rc;
    const auto wall0 = std::chrono::steady_clock::now();
    if (!on_device) {  // large host vectors: copies, permutations and tiles overlapped band by band
        bool used = false;
        rc = fvm_rhs_pipelined(h, t, u, du, &used);
        if (rc || used) {
            if (!rc) fvm_pipe_report(h, 0, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
            return rc;
        }
    }
    const size_t bytes = sizeof(double) * h->N * h->neq;
    const double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, h->d_u, true))) return rc;
    if ((rc = fvm_apply_rhs(h, t, h->d_u, h->d_du))) return rc;
    if (on_device) {
        if ((rc = fvm_launch_permute(h, h->d_du, du, false))) return rc;
    } else {
        if ((rc = fvm_launch_permute(h, h->d_du, h->d_io, false))) return rc;
        FVM_CUDA(h, cudaMemcpyAsync(du, h->d_io, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!on_device) fvm_pipe_report(h, 0, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
    return FVM_OK;
}

extern "C" int32_t fvm_apply_dirichlet_native(fvm_handle h, double t, double* u) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u, "fvm_apply_dirichlet_native: null argument");
    return fvm_launch_dirichlet(h, t, u);
}

extern "C" int32_t fvm_apply_dirichlet(fvm_handle h, double t, double* u, int32_t on_device) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u, "fvm_apply_dirichlet: null argument");
    int32_t rc = fvm_ensure_state(h);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * h->N * h->neq;
    double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, h->d_u, true))) return rc;
    if ((rc = fvm_launch_dirichlet(h, t, h->d_u))) return rc;
    if ((rc = fvm_launch_permute(h, h->d_u, src, false))) return rc;
    if (!on_device) FVM_CUDA(h, cudaMemcpyAsync(u, h->d_io, bytes, cudaMemcpyDeviceToHost, h->stream));
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FVM_OK;
}

extern "C" int32_t fvm_to_native(fvm_handle h, const double* v_caller, double* v_native) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, v_caller && v_native && v_caller != v_native, "fvm_to_native: bad arguments");
    return fvm_launch_permute(h, v_caller, v_native, true);
}

extern "C" int32_t fvm_from_native(fvm_handle h, const double* v_native, double* v_caller) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, v_caller && v_native && v_caller != v_native, "fvm_from_native: bad arguments");
    return fvm_launch_permute(h, v_native, v_caller, false);
}

// ---- optional CUDA-event timing of the dominant kernel, on the launching stream ---------------
void fvm_prof_begin(fvm_ctx* h) {
    if (!h->profiling || h->prof_used + 2 > (int64_t)h->prof_ev.size()) return;
    cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
}
void fvm_prof_end(fvm_ctx* h) {
    if (!h->profiling || h->prof_used + 2 > (int64_t)h->prof_ev.size()) return;
    cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream);
    h->prof_used += 2;
}

extern "C" int32_t fvm_set_profiling(fvm_handle h, int32_t max_launches) {
    if (!h) return FVM_ERR_ARG;
    FVM_CUDA(h, cudaSetDevice(h->device));
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    h->prof_ev.clear();
    h->prof_used = 0;
    h->profiling = max_launches > 0;
    for (int i = 0; i < 2 * max_launches; ++i) {
        cudaEvent_t e;
        FVM_CUDA(h, cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    return FVM_OK;
}

extern "C" int32_t fvm_get_profile(fvm_handle h, double* total_ms, int64_t* launches) {
    if (!h || !total_ms || !launches) return FVM_ERR_ARG;
    FVM_CUDA(h, cudaSetDevice(h->device));
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (int64_t i = 0; i + 1 < h->prof_used; i += 2) {
        float ms = 0.f;
        FVM_CUDA(h, cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = h->prof_used / 2;
    h->prof_used = 0;
    return FVM_OK;
}

extern "C" int32_t fvm_stream_synchronize(fvm_handle h) {
    NEED_FINAL(h);
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FVM_OK;
}

extern "C" int32_t fvm_get_stream(fvm_handle h, void** stream) {
    if (!h || !stream) return FVM_ERR_ARG;
    *stream = (void*)h->stream;
    return FVM_OK;
}

extern "C" int32_t fvm_get_geometry(fvm_handle h, double* V, double* s9, double* mid6, double* nrm6, double* len3) {
    NEED_FINAL(h);
    const int64_t N = h->N, T = h->T;
    int32_t rc;
    if (V) {
        if ((rc = fvm_ensure_state(h))) return rc;
        // vol is a scalar-per-node array: permute with neq = 1
        double* tmp = nullptr;
        FVM_CUDA(h, cudaMalloc((void**)&tmp, sizeof(double) * N));
        permute_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(h->d_node_old_of_new, h->dm.vol, tmp, N, 1, 0);
        cudaError_t e = cudaMemcpyAsync(V, tmp, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaFree(tmp);
        FVM_CUDA(h, e);
    }
    if (s9 || mid6 || nrm6 || len3) {
        int32_t* d_perm = nullptr;
        FVM_CUDA(h, cudaMalloc((void**)&d_perm, sizeof(int32_t) * T));
        FVM_CUDA(h, cudaMemcpyAsync(d_perm, h->tri_old_of_new.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, h->stream));
        double* outs[4] = {s9, mid6, nrm6, len3};
        const int width[4] = {9, 6, 6, 3};
        for (int q = 0; q < 4; ++q) {  // one array at a time bounds the temporary device memory
            if (!outs[q]) continue;
            double* tmp = nullptr;
            cudaError_t e = cudaMalloc((void**)&tmp, sizeof(double) * T * width[q]);
            if (e != cudaSuccess) {
                cudaFree(d_perm);
                FVM_CUDA(h, e);
            }
            export_geometry_kernel<<<(unsigned)((T + 255) / 256), 256, 0, h->stream>>>(
                h->dm, h->d_tri_native, d_perm, q == 0 ? tmp : nullptr, q == 1 ? tmp : nullptr, q == 2 ? tmp : nullptr,
                q == 3 ? tmp : nullptr);
            e = cudaMemcpyAsync(outs[q], tmp, sizeof(double) * T * width[q], cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
            cudaFree(tmp);
            if (e != cudaSuccess) {
                cudaFree(d_perm);
                FVM_CUDA(h, e);
            }
        }
        cudaFree(d_perm);
    }
    return FVM_OK;
}
