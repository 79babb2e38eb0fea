// libfvmcuda: the linear-template operator path
// (/root/reference/src/specific_problems/abstract_templates.jl:73-325 and the five constructors).
//
//   * structural CSR pattern = jacobian_sparsity (src/solve.jl:56-77), built on the device from a
//     node -> triangle incidence list, native (tile-major) numbering;
//   * assembly by row ownership: one thread per node gathers the contributions of its incident
//     triangles in ascending triangle order -- no atomics, deterministic, and never a dense
//     n x n matrix (the reference allocates zeros(n,n), diffusion_equation.jl:82);
//   * fp64 CSR SpMV y = A x (+ b): a CTA streams the val/col arrays of a chunk of rows with
//     fully coalesced loads into shared memory, then sums each row in CSR order.
#include <algorithm>
#include <cstring>

#include <chrono>

#include <cub/device/device_scan.cuh>

#include "fvm_device.cuh"

#define MAX_ROW 64
#define SPMV_BLOCK 256
#define SPMV_ROWS 256

struct AsmEdge {          // one boundary edge for the template assembly (native ids)
    int32_t v[3];         // stored vertex triple of the adjacent triangle
    int32_t i, j;
    int32_t kind;         // fvm_edge_kind of the edge
    double Di, Dj;        // diffusion function at the two quarter points
    double ai, aj;        // Neumann function at the two quarter points
};

// ---- pattern ---------------------------------------------------------------------------------
// node -> triangle incidence lists (code = triangle << 2 | slot), see fvm_build_pattern
__global__ void n2t_count_kernel(int64_t n3, const int32_t* __restrict__ tri, int32_t* __restrict__ cnt) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n3) atomicAdd(cnt + tri[k], 1);
}
__global__ void n2t_fill_kernel(int64_t n3, const int32_t* __restrict__ tri, const int32_t* __restrict__ ptr, int32_t* __restrict__ fill,
                                int32_t* __restrict__ items) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n3) return;
    const int32_t g = tri[k];
    items[ptr[g] + atomicAdd(fill + g, 1)] = (int32_t)((k / 3) << 2 | (k % 3));
}
__global__ void n2t_sort_kernel(int n, const int32_t* __restrict__ ptr, int32_t* __restrict__ items) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int lo = ptr[g], hi = ptr[g + 1];
    for (int a = lo + 1; a < hi; ++a) {  // insertion sort: a node has ~6 incident triangles
        const int32_t v = items[a];
        int b = a - 1;
        while (b >= lo && items[b] > v) {
            items[b + 1] = items[b];
            --b;
        }
        items[b + 1] = v;
    }
}

__device__ __forceinline__ int row_neighbours(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                              const int32_t* __restrict__ tri, int g, int32_t* out) {
    int cnt = 0;
    out[cnt++] = g;
    for (int e = n2t_ptr[g]; e < n2t_ptr[g + 1]; ++e) {
        const int code = n2t[e];
        const int64_t t = code >> 2;
        const int slot = code & 3;
        for (int q = 1; q <= 2; ++q) {
            const int w = tri[3 * t + (slot + q) % 3];
            int pos = 0;
            while (pos < cnt && out[pos] < w) ++pos;
            if (pos < cnt && out[pos] == w) continue;
            if (cnt >= MAX_ROW) return -1;
            for (int k = cnt; k > pos; --k) out[k] = out[k - 1];
            out[pos] = w;
            ++cnt;
        }
    }
    return cnt;
}

__global__ void pattern_count_kernel(int n, const int32_t* n2t_ptr, const int32_t* n2t, const int32_t* tri, int32_t* rowlen) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    int32_t nb[MAX_ROW];
    rowlen[g] = row_neighbours(n2t_ptr, n2t, tri, g, nb);
}

__global__ void pattern_fill_kernel(int n, const int32_t* n2t_ptr, const int32_t* n2t, const int32_t* tri,
                                    const int32_t* rowptr, int32_t* col) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    int32_t nb[MAX_ROW];
    const int cnt = row_neighbours(n2t_ptr, n2t, tri, g, nb);
    for (int k = 0; k < cnt; ++k) col[rowptr[g] + k] = nb[k];
}

// ---- assembly --------------------------------------------------------------------------------
struct AsmArgs {
    int32_t template_id;
    int32_t quirks;
    double d_const;
    const double* dtab;        // [T][3] native triangle order, or null
    const double* node_value;  // [N] native, or null
    const double* source;      // [N] native, or null
    const int32_t* n2t_ptr;
    const int32_t* n2t;
    const int32_t* tri;
    const int32_t* rowptr;
    const int32_t* col;
    double* val;
    double* b;
    double* rowscale;
};

__device__ __forceinline__ int find_col(const int32_t* __restrict__ col, int beg, int len, int c) {
    for (int k = 0; k < len; ++k)
        if (col[beg + k] == c) return k;
    return 0;
}

// triangle_contributions!(A, ...) (abstract_templates.jl:73-99) gathered by row, then the per-node
// fix-ups of the five constructors.
__global__ void assemble_rows_kernel(const DevMesh m, const AsmArgs a) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m.n_nodes) return;
    const int beg = a.rowptr[g], len = a.rowptr[g + 1] - beg;
    double acc[MAX_ROW];
    for (int k = 0; k < len; ++k) acc[k] = 0.0;
    const int dpos = find_col(a.col, beg, len, g);
    const bool vertex = g < m.n_vertices;
    const uint8_t kind = vertex ? m.kind[g] : (uint8_t)3;
    const bool steady = a.template_id >= FVM_TPL_MEAN_EXIT_TIME;
    double bval = 0.0;
    if (!vertex) {  // fix_missing_vertices!, abstract_templates.jl:317-325
        acc[dpos] = 1.0;
    } else if (kind == FVM_NODE_FREE) {
        const double V = m.vol[g];
        for (int e = a.n2t_ptr[g]; e < a.n2t_ptr[g + 1]; ++e) {
            const int code = a.n2t[e];
            const int64_t t = code >> 2;
            const int p = code & 3;
            const int v[3] = {a.tri[3 * t], a.tri[3 * t + 1], a.tri[3 * t + 2]};
            TriGeom G;
            tri_geometry<true>(m.xy[2 * (size_t)v[0]], m.xy[2 * (size_t)v[0] + 1], m.xy[2 * (size_t)v[1]],
                               m.xy[2 * (size_t)v[1] + 1], m.xy[2 * (size_t)v[2]], m.xy[2 * (size_t)v[2] + 1], G, nullptr);
            int pos[3];
            for (int k = 0; k < 3; ++k) pos[k] = find_col(a.col, beg, len, v[k]);
            // this node is e1 of cv-edge p (sign +) and e2 of cv-edge (p+2)%3 (sign -); keep the
            // reference's edge order 1,2,3 so the additions happen in the same sequence
            for (int ed = 0; ed < 3; ++ed) {
                const bool plus = (ed == p), minus = (ed == (p + 2) % 3);
                if (!plus && !minus) continue;
                const double l = __dsqrt_rn(__dadd_rn(__dmul_rn(G.ex[ed], G.ex[ed]), __dmul_rn(G.ey[ed], G.ey[ed])));
                const double nx = __ddiv_rn(G.ey[ed], l), ny = __ddiv_rn(-G.ex[ed], l);
                const double D = a.dtab ? a.dtab[3 * t + ed] : a.d_const;
                const double Dl = __dmul_rn(D, l);
                for (int k = 0; k < 3; ++k) {
                    const double a123 = __dmul_rn(Dl, __dadd_rn(__dmul_rn(G.s[k], nx), __dmul_rn(G.s[3 + k], ny)));
                    const double c = __ddiv_rn(a123, V);
                    acc[pos[k]] = plus ? __dadd_rn(acc[pos[k]], c) : __dsub_rn(acc[pos[k]], c);
                }
            }
        }
        if (a.template_id == FVM_TPL_LINEAR_REACTION_DIFFUSION && a.source)  // linear_source_contributions!
            acc[dpos] = __dadd_rn(acc[dpos], a.source[g]);
        if (a.template_id == FVM_TPL_POISSON && a.source) bval = a.source[g];  // create_rhs_b
        if (a.template_id == FVM_TPL_MEAN_EXIT_TIME) bval = -1.0;             // create_met_b!
    } else if (kind == FVM_NODE_DIRICHLET) {
        if (steady) {  // apply_steady_dirichlet_conditions! / create_met_b!
            acc[dpos] = 1.0;
            bval = (a.template_id == FVM_TPL_MEAN_EXIT_TIME || !a.node_value) ? 0.0 : a.node_value[g];
        }
    } else if (kind == FVM_NODE_DUDT) {  // apply_dudt_conditions!, abstract_templates.jl:127-135
        bval = a.node_value ? a.node_value[g] : 0.0;
    }
    for (int k = 0; k < len; ++k) a.val[beg + k] = acc[k];
    a.b[g] = bval;
    // symmetrising row scale for PCG: -V on free rows (A = -(1/V) K), 1 on identity / frozen rows
    a.rowscale[g] = (vertex && kind == FVM_NODE_FREE) ? -m.vol[g] : 1.0;
}

// boundary_edge_contributions! (abstract_templates.jl:149-191, 237-267) by boundary node
__global__ void assemble_boundary_kernel(const DevMesh m, const AsmArgs a, const AsmEdge* __restrict__ edges,
                                         const int32_t* __restrict__ bn_node, const int32_t* __restrict__ bn_ptr,
                                         const int32_t* __restrict__ bn_items, const int n_bn) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_bn) return;
    const int g = bn_node[q];
    if (m.kind[g] != FVM_NODE_FREE) return;  // rows of conditioned nodes are never touched
    const int beg = a.rowptr[g], len = a.rowptr[g + 1] - beg;
    for (int it = bn_ptr[q]; it < bn_ptr[q + 1]; ++it) {
        const AsmEdge E = edges[bn_items[it] >> 1];
        const int role = bn_items[it] & 1;  // 0: this node is i, 1: this node is j
        const double px = m.xy[2 * (size_t)E.i], py = m.xy[2 * (size_t)E.i + 1];
        const double qx = m.xy[2 * (size_t)E.j], qy = m.xy[2 * (size_t)E.j + 1];
        // get_boundary_cv_components, control_volumes.jl:41-56
        const double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py);
        const double lij = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        const double nx = __ddiv_rn(dy, lij), ny = __ddiv_rn(-dx, lij);
        const double mijx = __dmul_rn(__dadd_rn(px, qx), 0.5), mijy = __dmul_rn(__dadd_rn(py, qy), 0.5);
        const double hx = __dsub_rn(mijx, px), hy = __dsub_rn(mijy, py);
        const double l = __dsqrt_rn(__dadd_rn(__dmul_rn(hx, hx), __dmul_rn(hy, hy)));
        if (E.kind == FVM_EDGE_NEUMANN) {
            const double D = role ? E.Dj : E.Di, av = role ? E.aj : E.ai;
            a.b[g] = __dadd_rn(a.b[g], __ddiv_rn(__dmul_rn(__dmul_rn(D, av), l), m.vol[g]));
        } else {
            TriGeom G;
            tri_geometry<true>(m.xy[2 * (size_t)E.v[0]], m.xy[2 * (size_t)E.v[0] + 1], m.xy[2 * (size_t)E.v[1]],
                               m.xy[2 * (size_t)E.v[1] + 1], m.xy[2 * (size_t)E.v[2]], m.xy[2 * (size_t)E.v[2] + 1], G, nullptr);
            const double D = role ? E.Dj : E.Di;
            // abstract_templates.jl:262 divides the j row by V_i
            const double V = (role && a.quirks) ? m.vol[E.i] : m.vol[g];
            const double Dl = __dmul_rn(D, l);
            for (int k = 0; k < 3; ++k) {
                const double c = __ddiv_rn(__dmul_rn(Dl, __dadd_rn(__dmul_rn(G.s[k], nx), __dmul_rn(G.s[3 + k], ny))), V);
                const int pos = find_col(a.col, beg, len, E.v[k]);
                a.val[beg + pos] = __dadd_rn(a.val[beg + pos], c);
            }
        }
    }
}

__global__ void jacobi_kernel(int n, const int32_t* rowptr, const int32_t* col, const double* val, const double* rowscale,
                              double* diag_inv) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double d = 0.0;
    for (int q = rowptr[g]; q < rowptr[g + 1]; ++q)
        if (col[q] == g) d = val[q];
    d *= rowscale[g];
    diag_inv[g] = d != 0.0 ? 1.0 / d : 1.0;
}

// ---- SpMV ------------------------------------------------------------------------------------
// y = A x (+ b) [* rowscale].  A CTA owns SPMV_ROWS consecutive rows: their val/col entries are one
// contiguous span of the CSR arrays, streamed with coalesced loads; products land in shared memory
// and each row is then summed in CSR order by one thread (deterministic).
template <bool ADD_B, bool SCALE>
__global__ void __launch_bounds__(SPMV_BLOCK)
    spmv_kernel(const int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ b, const double* __restrict__ rowscale,
                const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ double prod[];
    __shared__ int rp[SPMV_ROWS + 1];
    const int r0 = blockIdx.x * SPMV_ROWS;
    const int nr = min(SPMV_ROWS, n - r0);
    for (int i = threadIdx.x; i <= nr; i += SPMV_BLOCK) rp[i] = rowptr[r0 + i];
    __syncthreads();
    const int base = rp[0], cnt = rp[nr] - base;
    for (int q = threadIdx.x; q < cnt; q += SPMV_BLOCK) prod[q] = val[base + q] * __ldg(x + col[base + q]);
    __syncthreads();
    for (int i = threadIdx.x; i < nr; i += SPMV_BLOCK) {
        double s = 0.0;
        for (int q = rp[i] - base; q < rp[i + 1] - base; ++q) s += prod[q];
        if (ADD_B) s += b[r0 + i];
        if (SCALE) s *= rowscale[r0 + i];
        y[r0 + i] = s;
    }
}

// Warp-granular CSR SpMV: a warp owns 32 consecutive rows whose val/col entries are one contiguous
// span; lanes stream the span coalesced (all lanes busy, ~7 independent loads each), park the
// products in a warp-private shared-memory segment, then every lane sums its own row in CSR order.
// No block-level barrier: warps never wait for each other.
template <bool ADD_B, bool SCALE>
__global__ void __launch_bounds__(SPMV_BLOCK)
    spmv_warp_kernel(const int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const double* __restrict__ val, const double* __restrict__ b, const double* __restrict__ rowscale,
                     const double* __restrict__ x, double* __restrict__ y, const int cap) {
    extern __shared__ double prod_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* prod = prod_all + (size_t)warp * cap;
    const int r0 = (blockIdx.x * (SPMV_BLOCK / 32) + warp) * 32;
    if (r0 >= n) return;
    const int row = r0 + lane;
    const int rp = __ldg(rowptr + min(row, n));
    const int rp_next = __ldg(rowptr + min(row + 1, n));
    const int base = __shfl_sync(0xffffffffu, rp, 0);
    const int end = __shfl_sync(0xffffffffu, rp_next, 31);
    const int cnt = end - base;
#pragma unroll 8
    for (int q = lane; q < cnt; q += 32) prod[q] = __ldg(val + base + q) * __ldg(x + __ldg(col + base + q));
    __syncwarp();
    if (row < n) {
        double s = 0.0;
        for (int q = rp - base; q < rp_next - base; ++q) s += prod[q];
        if (ADD_B) s += b[row];
        if (SCALE) s *= rowscale[row];
        y[row] = s;
    }
}

// ---- tile-local SpMV ---------------------------------------------------------------------------
// Reuses the RHS tiling: a CTA owns the interior rows of one tile.  x of the tile's local nodes is
// staged in shared memory (own range coalesced, external interface nodes gathered), columns are
// 16-bit tile-local ids: 10 B per nonzero instead of 12, and no global gather of x at all.
// fixed-shape sum over the CTA (warp shuffles, then the warps in ascending order): deterministic
template <int THREADS>
__device__ __forceinline__ double cta_sum(double v) {
    __shared__ double sh[THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += sh[w];
    return s;
}

// interface rows and points that are not vertices: sliced ELL with global columns, one lane per row.  Slice `sl` of the
// tail (or nothing when !live); with FUSE == 2 every thread of the CTA must call it (block reduction inside).
struct TailArgs {
    int32_t n_rows = 0;
    const int32_t* rows = nullptr;
    const int32_t* sptr = nullptr;
    const int32_t* scol = nullptr;
    const double* sval = nullptr;
    int32_t n_blocks = 0;  // CTAs of a merged launch that work on tail slices (SPMV_BLOCK / 32 slices each)
};
template <bool ADD_B, bool SCALE, int FUSE, int THREADS>
__device__ __forceinline__ void tail_slice(const TailArgs& ta, const int sl, const bool live, const double* __restrict__ b,
                                           const double* __restrict__ rowscale, const double* __restrict__ x, double* __restrict__ y,
                                           const SpmvFuse& f, const int dot_slot) {
    const int lane = threadIdx.x & 31;
    double dot = 0.0;
    if (live) {
        const int p0 = __ldg(ta.sptr + sl), len = (__ldg(ta.sptr + sl + 1) - p0) >> 5;
        const int k = sl * 32 + lane;
        const int g = k < ta.n_rows ? ta.rows[k] : -1;
        double acc = 0.0;
#pragma unroll 4
        for (int q = 0; q < len; ++q) acc += __ldg(ta.sval + p0 + q * 32 + lane) * x[__ldg(ta.scol + p0 + q * 32 + lane)];
        if (g >= 0) {
            if (ADD_B) acc += b[g];
            if (SCALE) acc *= rowscale[g];
            y[g] = acc;
            if constexpr (FUSE == 2) dot = x[g] * acc;
        }
    }
    if constexpr (FUSE == 2) {
        const double tot = cta_sum<THREADS>(dot);
        if (threadIdx.x == 0) f.dotpart[dot_slot] = tot;
    }
}

// One launch applies the whole operator: the first ta.n_blocks CTAs take the tail slices (they are gather-latency bound and
// independent of the tiles, so they start first and overlap with the bandwidth-bound tile CTAs that fill the other slots),
// every other CTA owns the interior rows of one tile.
template <bool ADD_B, bool SCALE, int FUSE>
__global__ void __launch_bounds__(SPMV_BLOCK)
    spmv_tile_kernel(const DevMesh m, const int32_t* __restrict__ tile_slice0, const int32_t* __restrict__ sell_ptr,
                     const uint16_t* __restrict__ sell_col, const double* __restrict__ sell_val, const double* __restrict__ b,
                     const double* __restrict__ rowscale, const double* __restrict__ x, double* __restrict__ y,
                     const int32_t* __restrict__ tile_list, const int tile_off, const SpmvFuse f, const TailArgs ta) {
    extern __shared__ double x_s[];  // [max_nloc]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x < ta.n_blocks) {
        if constexpr (FUSE == 2) {
            if (f.sc[SC_DONE] != 0.0) return;
        }
        const int sl = blockIdx.x * (SPMV_BLOCK / 32) + warp;
        tail_slice<ADD_B, SCALE, FUSE, SPMV_BLOCK>(ta, sl, sl * 32 < ta.n_rows, b, rowscale, x, y, f, m.n_tiles + blockIdx.x);
        return;
    }
    const int bid = blockIdx.x - ta.n_blocks;
    const int tile = tile_list ? tile_list[bid + tile_off] : bid;
    const int4 m0 = __ldg(m.tile_meta + 2 * tile), m1 = __ldg(m.tile_meta + 2 * tile + 1);
    const int node0 = m0.x, nint = m0.y, nown = m0.z, nloc = m0.w, ext0 = m1.x;
    if constexpr (FUSE == 2) {
        if (f.sc[SC_DONE] != 0.0) return;
        if (nint == 0) {  // uniform: no row of this tile is computed here; its partial still has to exist
            if (tid == 0) f.dotpart[tile] = 0.0;
            return;
        }
    } else {
        if (nint == 0) return;
    }
    for (int i = tid; i < nown; i += SPMV_BLOCK) x_s[i] = x[node0 + i];
    for (int k = tid; k < nloc - nown; k += SPMV_BLOCK) x_s[nown + k] = x[m.ext_ids[ext0 + k]];
    const int s0 = __ldg(tile_slice0 + tile), s1 = __ldg(tile_slice0 + tile + 1);
    __syncthreads();
    double dot = 0.0;
    // sliced ELL: entry k of the 32 rows of a slice is contiguous -> every load below is coalesced,
    // all 2*len loads of a row are independent, and the row is summed in CSR order in a register
    for (int sl = s0 + warp; sl < s1; sl += SPMV_BLOCK / 32) {
        const int p0 = __ldg(sell_ptr + sl), len = (__ldg(sell_ptr + sl + 1) - p0) >> 5;
        const double* __restrict__ v = sell_val + p0 + lane;
        const uint16_t* __restrict__ c = sell_col + p0 + lane;
        double acc = 0.0;
        int k = 0;
        for (; k + 8 <= len; k += 8) {
            double vv[8];
            uint16_t cc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                vv[q] = __ldg(v + (k + q) * 32);
                cc[q] = __ldg(c + (k + q) * 32);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += vv[q] * x_s[cc[q]];
        }
        {
            double vv[8];
            uint16_t cc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool on = k + q < len;
                vv[q] = on ? __ldg(v + (k + q) * 32) : 0.0;
                cc[q] = on ? __ldg(c + (k + q) * 32) : (uint16_t)0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (k + q < len) acc += vv[q] * x_s[cc[q]];
        }
        const int l = (sl - s0) * 32 + lane;
        if (l < nint) {
            if (ADD_B) acc += b[node0 + l];
            if (SCALE) acc *= rowscale[node0 + l];
            y[node0 + l] = acc;
            if constexpr (FUSE == 2) dot += x_s[l] * acc;
        }
    }
    if constexpr (FUSE == 2) {
        const double tot = cta_sum<SPMV_BLOCK>(dot);
        if (tid == 0) f.dotpart[tile] = tot;
    }
}

// fills the sliced-ELL arrays of the interior rows from the CSR arrays (cols once, vals per assembly)
__global__ void sell_pack_kernel(const DevMesh m, const int32_t* __restrict__ tile_slice0, const int32_t* __restrict__ sell_ptr,
                                 const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ col16,
                                 const double* __restrict__ val, uint16_t* __restrict__ sell_col, double* __restrict__ sell_val) {
    const int tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int4 m0 = m.tile_meta[2 * tile];
    const int node0 = m0.x, nint = m0.y;
    const int s0 = tile_slice0[tile], s1 = tile_slice0[tile + 1];
    for (int sl = s0 + warp; sl < s1; sl += blockDim.x / 32) {
        const int p0 = sell_ptr[sl], len = (sell_ptr[sl + 1] - p0) >> 5;
        const int l = (sl - s0) * 32 + lane;
        const int rb = l < nint ? rowptr[node0 + l] : 0, rl = l < nint ? rowptr[node0 + l + 1] - rb : 0;
        for (int k = 0; k < len; ++k) {
            if (sell_col) sell_col[p0 + k * 32 + lane] = k < rl ? col16[rb + k] : (uint16_t)0;
            if (sell_val) sell_val[p0 + k * 32 + lane] = k < rl ? val[rb + k] : 0.0;
        }
    }
}

// the tail slices on their own (host-buffer pipeline: explicit slice lists; sharded runs without a tile path)
template <bool ADD_B, bool SCALE>
__global__ void __launch_bounds__(128)
    spmv_rows_kernel(const TailArgs ta, const double* __restrict__ b, const double* __restrict__ rowscale, const double* __restrict__ x,
                     double* __restrict__ y, const int32_t* __restrict__ slice_list, const int list_off, const int list_count) {
    int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    bool live = true;
    if (slice_list) {  // explicit subset of the slices
        if (sl >= list_count) live = false;
        else sl = slice_list[list_off + sl];
    }
    if (!live || sl * 32 >= ta.n_rows) return;
    tail_slice<ADD_B, SCALE, 0, 128>(ta, sl, true, b, rowscale, x, y, SpmvFuse{}, 0);
}

__global__ void tsell_pack_kernel(const int n_rows, const int32_t* __restrict__ rows, const int32_t* __restrict__ sptr,
                                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  const double* __restrict__ val, int32_t* __restrict__ scol, double* __restrict__ sval) {
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (sl * 32 >= n_rows) return;
    const int p0 = sptr[sl], len = (sptr[sl + 1] - p0) >> 5;
    const int k = sl * 32 + lane;
    const int g = k < n_rows ? rows[k] : -1;
    const int rb = g >= 0 ? rowptr[g] : 0, rl = g >= 0 ? rowptr[g + 1] - rb : 0;
    for (int q = 0; q < len; ++q) {
        if (scol) scol[p0 + q * 32 + lane] = q < rl ? col[rb + q] : (g >= 0 ? g : 0);
        if (sval) sval[p0 + q * 32 + lane] = q < rl ? val[rb + q] : 0.0;
    }
}

// tile-local column ids of the entries of interior rows (one-off, at pattern build)
__global__ void col16_kernel(const DevMesh m, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             uint16_t* __restrict__ col16) {
    const int tile = blockIdx.x;
    const int4 m0 = m.tile_meta[2 * tile], m1 = m.tile_meta[2 * tile + 1];
    const int node0 = m0.x, nint = m0.y, nown = m0.z, nloc = m0.w, ext0 = m1.x;
    if (nint == 0) return;
    const int base = rowptr[node0], end = rowptr[node0 + nint];
    for (int q = base + threadIdx.x; q < end; q += blockDim.x) {
        const int g = col[q];
        int l = g - node0;
        if (l < 0 || l >= nown) {
            l = 0xffff;
            for (int k = 0; k < nloc - nown; ++k)
                if (m.ext_ids[ext0 + k] == g) {
                    l = nown + k;
                    break;
                }
        }
        col16[q] = (uint16_t)l;
    }
}

static const SpmvFuse kNoFuse{};

template <bool ADD_B, bool SCALE, int FUSE>
static int32_t launch_spmv_t(fvm_ctx* h, const double* x, double* y, int part, const SpmvFuse& f) {
    Csr& c = h->csr;
    if (c.use_tile_spmv) {
        int grid = h->dm.n_tiles, off = 0;
        const int32_t* list = nullptr;
        if (part == 1) {
            list = h->d_tile_order;
            grid = h->n_tiles_indep;
        } else if (part == 2) {
            list = h->d_tile_order;
            off = h->n_tiles_indep;
            grid = h->dm.n_tiles - h->n_tiles_indep;
        } else if (part == 4) {  // explicit tile list (host-buffer pipeline)
            list = h->pipe_list;
            off = h->pipe_off;
            grid = h->pipe_count;
        }
        cudaStream_t st = h->launch_stream;
        // the tail slices ride in the launch that may read the whole input: everything (0) or the halo-dependent part (2)
        TailArgs ta;
        ta.n_rows = c.n_tail;
        ta.rows = c.tail_rows;
        ta.sptr = c.tsell_ptr;
        ta.scol = c.tsell_col;
        ta.sval = c.tsell_val;
        ta.n_blocks = (c.n_tail > 0 && (part == 0 || part == 2)) ? (c.n_tslices + SPMV_BLOCK / 32 - 1) / (SPMV_BLOCK / 32) : 0;
        if (grid + ta.n_blocks > 0 && part != 3) {
            auto kern = spmv_tile_kernel<ADD_B, SCALE, FUSE>;
            // the opt-in is per function and per device, not per handle: always the largest size the tile path accepts
            int32_t& configured = h->smem_configured[(const void*)kern];
            if (c.tile_smem > 48 * 1024 && configured == 0) {
                FVM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                configured = 1;
            }
            if (st == h->stream) fvm_prof_begin(h);
            kern<<<grid + ta.n_blocks, SPMV_BLOCK, c.tile_smem, st>>>(h->dm, c.tile_slice0, c.sell_ptr, c.sell_col, c.sell_val, c.b, c.rowscale,
                                                                      x, y, list, off, f, ta);
            if (st == h->stream) fvm_prof_end(h);
        }
    } else if (FUSE != 0) {
        return fvm_fail(h, FVM_ERR_STATE, "fused SpMV needs the tile kernels");
    } else if (part == 1 || part == 2) {
        return FVM_OK;  // the generic kernels have no tile split: everything runs in part 3
    } else if (c.use_tile_spmv == 0 && c.chunk_rows > 0 && getenv("FVM_SPMV_BLOCK")) {
        const int grid = (c.n + SPMV_ROWS - 1) / SPMV_ROWS;
        fvm_prof_begin(h);
        spmv_kernel<ADD_B, SCALE><<<grid, SPMV_BLOCK, c.chunk_smem, h->stream>>>(c.n, c.rowptr, c.col, c.val, c.b, c.rowscale, x, y);
        fvm_prof_end(h);
    } else {
        const int cap = 32 * c.max_row;
        const int grid = (c.n + SPMV_BLOCK - 1) / SPMV_BLOCK;
        fvm_prof_begin(h);
        spmv_warp_kernel<ADD_B, SCALE><<<grid, SPMV_BLOCK, sizeof(double) * cap * (SPMV_BLOCK / 32), h->stream>>>(
            c.n, c.rowptr, c.col, c.val, c.b, c.rowscale, x, y, cap);
        fvm_prof_end(h);
    }
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

template <int FUSE>
static int32_t launch_spmv_f(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale, int part, const SpmvFuse& f) {
    if (add_b && !scale) return launch_spmv_t<true, false, FUSE>(h, x, y, part, f);
    if (!add_b && !scale) return launch_spmv_t<false, false, FUSE>(h, x, y, part, f);
    if (add_b && scale) return launch_spmv_t<true, true, FUSE>(h, x, y, part, f);
    return launch_spmv_t<false, true, FUSE>(h, x, y, part, f);
}

static int32_t launch_spmv_any(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale, int part, const SpmvFuse& f) {
    if (f.kind == 2) return launch_spmv_f<2>(h, x, y, add_b, scale, part, f);
    return launch_spmv_f<0>(h, x, y, add_b, scale, part, f);
}

int32_t fvm_launch_spmv_part(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale, int part) {
    return launch_spmv_any(h, x, y, add_b, scale, part, kNoFuse);
}

int32_t fvm_launch_spmv(fvm_ctx* h, const double* x, double* y, bool add_b, bool scale) {
    return fvm_launch_spmv_part(h, x, y, add_b, scale, 0);
}

// tail-row slices list[off .. off+count) (interface rows and points that are not vertices)
int32_t fvm_launch_spmv_tail_list(fvm_ctx* h, const double* x, double* y, bool add_b, const int32_t* list, int off, int count) {
    Csr& c = h->csr;
    if (count <= 0) return FVM_OK;
    const int grid = (count * 32 + 127) / 128;
    TailArgs ta;
    ta.n_rows = c.n_tail;
    ta.rows = c.tail_rows;
    ta.sptr = c.tsell_ptr;
    ta.scol = c.tsell_col;
    ta.sval = c.tsell_val;
    if (add_b) spmv_rows_kernel<true, false><<<grid, 128, 0, h->launch_stream>>>(ta, c.b, c.rowscale, x, y, list, off, count);
    else spmv_rows_kernel<false, false><<<grid, 128, 0, h->launch_stream>>>(ta, c.b, c.rowscale, x, y, list, off, count);
    FVM_CUDA(h, cudaGetLastError());
    return FVM_OK;
}

bool fvm_spmv_fusable(fvm_ctx* h) { return h->csr.assembled && h->csr.use_tile_spmv != 0; }

int32_t fvm_spmv_fused_partials(fvm_ctx* h) {
    return h->dm.n_tiles + (h->csr.n_tail > 0 ? (h->csr.n_tslices + SPMV_BLOCK / 32 - 1) / (SPMV_BLOCK / 32) : 0);
}

int32_t fvm_apply_spmv_fused(fvm_ctx* h, double* x, double* y, bool add_b, bool scale, const SpmvFuse& f) {
    if (!h->halo_ready) return launch_spmv_any(h, x, y, add_b, scale, 0, f);
    int32_t rc;
    if (!h->overlap) {
        if ((rc = fvm_halo_exchange(h, x))) return rc;
        return launch_spmv_any(h, x, y, add_b, scale, 0, f);
    }
    if ((rc = fvm_halo_begin(h, x))) return rc;
    h->launch_stream = h->comm_stream;  // halo-dependent tiles right behind the unpack, on the communication stream
    rc = launch_spmv_any(h, x, y, add_b, scale, 2, f);
    h->launch_stream = h->stream;
    if (rc) return rc;
    if ((rc = fvm_halo_done(h))) return rc;
    if ((rc = launch_spmv_any(h, x, y, add_b, scale, 1, f))) return rc;
    if ((rc = fvm_halo_wait(h))) return rc;
    return launch_spmv_any(h, x, y, add_b, scale, 3, f);
}

int32_t fvm_apply_spmv(fvm_ctx* h, double* x, double* y, bool add_b, bool scale) {
    return fvm_apply_spmv_fused(h, x, y, add_b, scale, kNoFuse);
}

// ---- host side -------------------------------------------------------------------------------
#define NEED_FINAL(h)                                                                       \
    do {                                                                                    \
        if (!(h)) return FVM_ERR_ARG;                                                       \
        if (!(h)->finalized) return fvm_fail((h), FVM_ERR_STATE, "call fvm_finalize first"); \
        FVM_CUDA(h, cudaSetDevice((h)->device));                                            \
    } while (0)

int32_t fvm_build_pattern(fvm_ctx* h) {
    Csr& c = h->csr;
    if (c.pattern) return FVM_OK;
    const int64_t N = h->N, T = h->T;
    // node -> incident (triangle, slot) lists, built on the device from the native triangle table: count, exclusive scan,
    // fill (integer atomics give each entry a position), then every node's few entries are sorted ascending -- the order
    // the assembly sums in, so A does not depend on the scheduling of the fill (round 1 built these lists with serial
    // host loops over 3T random accesses and uploaded 0.5 GB: ~1 s at 4096^2)
    int32_t rc;
    if ((rc = fvm_dev_alloc(h, &c.n2t_ptr, (size_t)N + 1))) return rc;
    if ((rc = fvm_dev_alloc(h, &c.n2t, (size_t)3 * T))) return rc;
    {
        int32_t* cnt = nullptr;
        void* tmp = nullptr;
        size_t tmp_bytes = 0;
        FVM_CUDA(h, cudaMalloc((void**)&cnt, sizeof(int32_t) * ((size_t)N + 1)));
        cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)N + 1), h->stream);
        const unsigned gt = (unsigned)((3 * T + 255) / 256), gn = (unsigned)((N + 255) / 256);
        if (e == cudaSuccess) {
            n2t_count_kernel<<<gt, 256, 0, h->stream>>>(3 * T, h->d_tri_native, cnt);
            e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, c.n2t_ptr, (int)(N + 1), h->stream);
        }
        if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
        if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, c.n2t_ptr, (int)(N + 1), h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)N + 1), h->stream);
        if (e == cudaSuccess) {
            n2t_fill_kernel<<<gt, 256, 0, h->stream>>>(3 * T, h->d_tri_native, c.n2t_ptr, cnt, c.n2t);
            n2t_sort_kernel<<<gn, 256, 0, h->stream>>>((int)N, c.n2t_ptr, c.n2t);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(cnt);
        if (tmp) cudaFree(tmp);
        FVM_CUDA(h, e);
    }
    int32_t* rowlen = nullptr;
    FVM_CUDA(h, cudaMalloc((void**)&rowlen, sizeof(int32_t) * N));
    const unsigned grid = (unsigned)((N + 127) / 128);
    pattern_count_kernel<<<grid, 128, 0, h->stream>>>((int)N, c.n2t_ptr, c.n2t, h->d_tri_native, rowlen);
    std::vector<int32_t> hl(N), rp(N + 1, 0);
    cudaError_t e = cudaMemcpyAsync(hl.data(), rowlen, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(rowlen);
    FVM_CUDA(h, e);
    int64_t nnz = 0;
    int32_t maxrow = 0, chunk = 0;
    for (int64_t g = 0; g < N; ++g) {
        if (hl[g] < 0) return fvm_fail(h, FVM_ERR_ARG, "fvm_assemble: a node has more than 63 neighbours");
        maxrow = std::max(maxrow, hl[g]);
        nnz += hl[g];
        if (nnz >= INT32_MAX) return fvm_fail(h, FVM_ERR_ARG, "fvm_assemble: nnz exceeds int32");
        rp[g + 1] = (int32_t)nnz;
    }
    for (int64_t r0 = 0; r0 < N; r0 += SPMV_ROWS) chunk = std::max(chunk, rp[std::min<int64_t>(N, r0 + SPMV_ROWS)] - rp[r0]);
    c.n = (int32_t)N;
    c.nnz = nnz;
    c.max_row = maxrow;
    c.chunk_rows = SPMV_ROWS;
    c.chunk_smem = (int32_t)(sizeof(double) * chunk);
    if (c.chunk_smem > 200 * 1024) return fvm_fail(h, FVM_ERR_ARG, "fvm_assemble: SpMV row chunk exceeds shared memory");
    if ((rc = fvm_dev_upload(h, &c.rowptr, rp))) return rc;
    if ((rc = fvm_dev_alloc(h, &c.col, (size_t)nnz))) return rc;
    if ((rc = fvm_dev_alloc(h, &c.val, (size_t)nnz + 8))) return rc;  // slack: quads are read whole
    if ((rc = fvm_dev_alloc(h, &c.b, (size_t)N))) return rc;
    if ((rc = fvm_dev_alloc(h, &c.rowscale, (size_t)N))) return rc;
    if ((rc = fvm_dev_alloc(h, &c.diag_inv, (size_t)N))) return rc;
    pattern_fill_kernel<<<grid, 128, 0, h->stream>>>((int)N, c.n2t_ptr, c.n2t, h->d_tri_native, c.rowptr, c.col);
    FVM_CUDA(h, cudaGetLastError());
    if (c.chunk_smem > 48 * 1024) {
        FVM_CUDA(h, cudaFuncSetAttribute(spmv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.chunk_smem));
        FVM_CUDA(h, cudaFuncSetAttribute(spmv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.chunk_smem));
        FVM_CUDA(h, cudaFuncSetAttribute(spmv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.chunk_smem));
        FVM_CUDA(h, cudaFuncSetAttribute(spmv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.chunk_smem));
    }
    {   // tile-local SpMV structures
        const int64_t n_tiles = h->dm.n_tiles;
        std::vector<int4> meta(2 * n_tiles);
        FVM_CUDA(h, cudaMemcpyAsync(meta.data(), h->dm.tile_meta, sizeof(int4) * 2 * n_tiles, cudaMemcpyDeviceToHost, h->stream));
        FVM_CUDA(h, cudaStreamSynchronize(h->stream));
        int32_t cap = 0, max_nint = 0;
        std::vector<int32_t> tail;
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int node0 = meta[2 * b].x, nint = meta[2 * b].y, nown = meta[2 * b].z;
            cap = std::max(cap, rp[node0 + nint] - rp[node0]);
            max_nint = std::max(max_nint, nint);
            for (int l = nint; l < nown; ++l) tail.push_back(node0 + l);
        }
        for (int64_t g = h->dm.n_vertices; g < N; ++g) tail.push_back((int32_t)g);
        c.n_tail = (int32_t)tail.size();
        c.tile_prod_cap = cap;
        c.tile_max_nint = max_nint;
        c.tile_smem = (int32_t)((sizeof(double) * (size_t)h->max_nloc + 15) & ~(size_t)15);
        // sliced-ELL layout of the interior rows
        std::vector<int32_t> slice0(n_tiles + 1, 0), sptr(1, 0);
        int64_t entries = 0;
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int node0 = meta[2 * b].x, nint = meta[2 * b].y;
            slice0[b] = (int32_t)(sptr.size() - 1);
            for (int l0 = 0; l0 < nint; l0 += 32) {
                int mx = 0;
                for (int l = l0; l < std::min(nint, l0 + 32); ++l) mx = std::max(mx, rp[node0 + l + 1] - rp[node0 + l]);
                entries += (int64_t)32 * mx;
                if (entries >= INT32_MAX) return fvm_fail(h, FVM_ERR_ARG, "fvm_assemble: sliced-ELL size exceeds int32");
                sptr.push_back((int32_t)entries);
            }
        }
        slice0[n_tiles] = (int32_t)(sptr.size() - 1);
        c.n_slices = slice0[n_tiles];
        c.sell_entries = entries;
        if ((rc = fvm_dev_upload(h, &c.tile_slice0, slice0))) return rc;
        if ((rc = fvm_dev_upload(h, &c.sell_ptr, sptr))) return rc;
        if ((rc = fvm_dev_alloc(h, &c.sell_val, (size_t)entries + 32))) return rc;
        if ((rc = fvm_dev_alloc(h, &c.sell_col, (size_t)entries + 32))) return rc;
        if ((rc = fvm_dev_upload(h, &c.tail_rows, tail))) return rc;
        {
            std::vector<int32_t> tp(1, 0);
            int64_t te = 0;
            for (size_t k0 = 0; k0 < tail.size(); k0 += 32) {
                int mx = 0;
                for (size_t k = k0; k < std::min(tail.size(), k0 + 32); ++k) mx = std::max(mx, rp[tail[k] + 1] - rp[tail[k]]);
                te += (int64_t)32 * mx;
                tp.push_back((int32_t)te);
            }
            c.n_tslices = (int32_t)tp.size() - 1;
            if ((rc = fvm_dev_upload(h, &c.tsell_ptr, tp))) return rc;
            if ((rc = fvm_dev_alloc(h, &c.tsell_val, (size_t)te + 32))) return rc;
            if ((rc = fvm_dev_alloc(h, &c.tsell_col, (size_t)te + 32))) return rc;
        }
        if ((rc = fvm_dev_alloc(h, &c.col16, (size_t)nnz + 8))) return rc;
        FVM_CUDA(h, cudaMemsetAsync(c.col16, 0, sizeof(uint16_t) * ((size_t)nnz + 8), h->stream));
        col16_kernel<<<(unsigned)n_tiles, 256, 0, h->stream>>>(h->dm, c.rowptr, c.col, c.col16);
        sell_pack_kernel<<<(unsigned)n_tiles, 256, 0, h->stream>>>(h->dm, c.tile_slice0, c.sell_ptr, c.rowptr, c.col16, nullptr,
                                                                   c.sell_col, nullptr);
        if (c.n_tail > 0)
            tsell_pack_kernel<<<(c.n_tslices * 32 + 127) / 128, 128, 0, h->stream>>>(c.n_tail, c.tail_rows, c.tsell_ptr, c.rowptr, c.col,
                                                                                      nullptr, c.tsell_col, nullptr);
        FVM_CUDA(h, cudaGetLastError());
        if (c.tile_smem > 200 * 1024) c.use_tile_spmv = 0;  // (the tile kernels opt in to large shared memory at first launch)
        if (const char* e = getenv("FVM_SPMV_GENERIC")) c.use_tile_spmv = (e[0] == '1') ? 0 : c.use_tile_spmv;
        if (sizeof(double) * 32 * maxrow * (SPMV_BLOCK / 32) > 48 * 1024) {
            const int wb = (int)(sizeof(double) * 32 * maxrow * (SPMV_BLOCK / 32));
            FVM_CUDA(h, cudaFuncSetAttribute(spmv_warp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb));
            FVM_CUDA(h, cudaFuncSetAttribute(spmv_warp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb));
            FVM_CUDA(h, cudaFuncSetAttribute(spmv_warp_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb));
            FVM_CUDA(h, cudaFuncSetAttribute(spmv_warp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb));
        }
    }
    c.pattern = true;
    h->stats[10] = nnz;
    h->stats[11] = maxrow;
    return FVM_OK;
}

template <class Tp>
static int32_t upload_native_nodes(fvm_ctx* h, const Tp* caller, Tp** dev, std::vector<void*>& temps) {
    *dev = nullptr;
    if (!caller) return FVM_OK;
    std::vector<Tp> nat(h->N);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < h->N; ++g) nat[g] = caller[h->node_old_of_new[g]];
    FVM_CUDA(h, cudaMalloc((void**)dev, sizeof(Tp) * h->N));
    temps.push_back(*dev);
    FVM_CUDA(h, cudaMemcpy(*dev, nat.data(), sizeof(Tp) * h->N, cudaMemcpyHostToDevice));
    return FVM_OK;
}

extern "C" int32_t fvm_assemble(fvm_handle h, int32_t template_id, double d_const, const double* d_cv_edge,
                                const double* d_bnd, const double* node_value, const double* edge_value,
                                const double* source, int32_t reference_quirks) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, h->neq == 1, "fvm_assemble: the linear templates are scalar problems (neq == 1)");
    FVM_REQUIRE(h, template_id >= FVM_TPL_DIFFUSION && template_id <= FVM_TPL_LAPLACE, "fvm_assemble: unknown template");
    const int64_t N = h->N, T = h->T, Eb = h->Eb;
    const bool steady = template_id >= FVM_TPL_MEAN_EXIT_TIME;
    bool has_dudt = false, has_constrained = false;
    for (int64_t i = 0; i < N; ++i) has_dudt = has_dudt || h->h_nkind[0][i] == FVM_NODE_DUDT;
    for (int64_t e = 0; e < Eb; ++e) has_constrained = has_constrained || h->h_ekind[0][e] == FVM_EDGE_CONSTRAINED;
    // poissons_equation.jl:69-70, laplaces_equation.jl:61-62, mean_exit_time.jl:66-69
    if (steady && has_dudt)
        return fvm_fail(h, FVM_ERR_ARG, template_id == FVM_TPL_MEAN_EXIT_TIME ? "MeanExitTimeProblem does not support Dudt nodes."
                                                                              : "PoissonsEquation does not support Dudt nodes.");
    if (template_id == FVM_TPL_MEAN_EXIT_TIME && has_constrained)
        return fvm_fail(h, FVM_ERR_ARG, "MeanExitTimeProblem does not support Constrained edges.");
    FVM_REQUIRE(h, d_cv_edge == nullptr || d_bnd != nullptr || Eb == 0 || template_id == FVM_TPL_MEAN_EXIT_TIME,
                "fvm_assemble: tabulated D needs the boundary quarter-point table too");
    int32_t rc = fvm_build_pattern(h);
    if (rc) return rc;
    Csr& c = h->csr;
    std::vector<void*> temps;
    auto cleanup = [&]() {
        for (void* p : temps) cudaFree(p);
    };
    AsmArgs a{};
    a.template_id = template_id;
    a.quirks = reference_quirks;
    a.d_const = d_const;
    a.n2t_ptr = c.n2t_ptr;
    a.n2t = c.n2t;
    a.tri = h->d_tri_native;
    a.rowptr = c.rowptr;
    a.col = c.col;
    a.val = c.val;
    a.b = c.b;
    a.rowscale = c.rowscale;
    double *d_nv = nullptr, *d_src = nullptr, *d_dt = nullptr;
    if ((rc = upload_native_nodes(h, node_value, &d_nv, temps)) || (rc = upload_native_nodes(h, source, &d_src, temps))) {
        cleanup();
        return rc;
    }
    a.node_value = d_nv;
    a.source = d_src;
    if (d_cv_edge) {
        std::vector<double> nat((size_t)3 * T);
#pragma omp parallel for schedule(static)
        for (int64_t nt = 0; nt < T; ++nt)
            for (int e = 0; e < 3; ++e) nat[3 * nt + e] = d_cv_edge[3 * (int64_t)h->tri_old_of_new[nt] + e];
        cudaError_t ce = cudaMalloc((void**)&d_dt, sizeof(double) * 3 * T);
        if (ce == cudaSuccess) {
            temps.push_back(d_dt);
            ce = cudaMemcpy(d_dt, nat.data(), sizeof(double) * 3 * T, cudaMemcpyHostToDevice);
        }
        if (ce != cudaSuccess) {
            cleanup();
            FVM_CUDA(h, ce);
        }
        a.dtab = d_dt;
    }
    assemble_rows_kernel<<<(unsigned)((N + 127) / 128), 128, 0, h->stream>>>(h->dm, a);
    if (template_id != FVM_TPL_MEAN_EXIT_TIME && Eb > 0) {  // MET skips the boundary-edge pass (mean_exit_time.jl:72-74)
        std::vector<AsmEdge> edges(Eb);
        std::vector<std::pair<int32_t, int32_t>> items;  // (native node, edge << 1 | role)
        items.reserve(2 * Eb);
        for (int64_t e = 0; e < Eb; ++e) {
            AsmEdge& E = edges[e];
            const int32_t* v = h->h_tri.data() + 3 * (int64_t)h->h_edge_tri[e];
            // _safe_get_triangle_props returns the stored rotation (utils.jl:1-14)
            for (int q = 0; q < 3; ++q) E.v[q] = h->node_new_of_old[v[q]];
            E.i = h->node_new_of_old[h->h_bedge[2 * e]];
            E.j = h->node_new_of_old[h->h_bedge[2 * e + 1]];
            E.kind = h->h_ekind[0][e];
            E.Di = d_bnd ? d_bnd[2 * e] : d_const;
            E.Dj = d_bnd ? d_bnd[2 * e + 1] : d_const;
            E.ai = edge_value ? edge_value[2 * e] : 0.0;
            E.aj = edge_value ? edge_value[2 * e + 1] : 0.0;
            items.push_back({E.i, (int32_t)(e << 1)});
            items.push_back({E.j, (int32_t)(e << 1 | 1)});
        }
        std::stable_sort(items.begin(), items.end(), [](auto& x, auto& y) { return x.first < y.first; });
        std::vector<int32_t> bn_node, bn_ptr, bn_items;
        for (size_t k = 0; k < items.size(); ++k) {
            if (k == 0 || items[k].first != items[k - 1].first) {
                bn_node.push_back(items[k].first);
                bn_ptr.push_back((int32_t)k);
            }
            bn_items.push_back(items[k].second);
        }
        bn_ptr.push_back((int32_t)items.size());
        AsmEdge* d_edges = nullptr;
        int32_t *d_bn = nullptr, *d_bp = nullptr, *d_bi = nullptr;
        cudaError_t ce = cudaMalloc((void**)&d_edges, sizeof(AsmEdge) * Eb);
        if (ce == cudaSuccess) temps.push_back(d_edges), ce = cudaMalloc((void**)&d_bn, sizeof(int32_t) * bn_node.size());
        if (ce == cudaSuccess) temps.push_back(d_bn), ce = cudaMalloc((void**)&d_bp, sizeof(int32_t) * bn_ptr.size());
        if (ce == cudaSuccess) temps.push_back(d_bp), ce = cudaMalloc((void**)&d_bi, sizeof(int32_t) * bn_items.size());
        if (ce == cudaSuccess) temps.push_back(d_bi);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_edges, edges.data(), sizeof(AsmEdge) * Eb, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_bn, bn_node.data(), sizeof(int32_t) * bn_node.size(), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_bp, bn_ptr.data(), sizeof(int32_t) * bn_ptr.size(), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_bi, bn_items.data(), sizeof(int32_t) * bn_items.size(), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            cleanup();
            FVM_CUDA(h, ce);
        }
        const int n_bn = (int)bn_node.size();
        assemble_boundary_kernel<<<(n_bn + 127) / 128, 128, 0, h->stream>>>(h->dm, a, d_edges, d_bn, d_bp, d_bi, n_bn);
    }
    sell_pack_kernel<<<(unsigned)h->dm.n_tiles, 256, 0, h->stream>>>(h->dm, c.tile_slice0, c.sell_ptr, c.rowptr, c.col16, c.val, nullptr,
                                                                     c.sell_val);
    if (c.n_tail > 0)
        tsell_pack_kernel<<<(c.n_tslices * 32 + 127) / 128, 128, 0, h->stream>>>(c.n_tail, c.tail_rows, c.tsell_ptr, c.rowptr, c.col, c.val,
                                                                                  nullptr, c.tsell_val);
    jacobi_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>((int)N, c.rowptr, c.col, c.val, c.rowscale, c.diag_inv);
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cleanup();
    FVM_CUDA(h, ce);
    c.assembled = true;
    c.template_id = template_id;
    return FVM_OK;
}

#define NEED_ASSEMBLED(h)                                                                             \
    do {                                                                                              \
        NEED_FINAL(h);                                                                                \
        if (!(h)->csr.assembled) return fvm_fail((h), FVM_ERR_STATE, "call fvm_assemble first");      \
    } while (0)

extern "C" int32_t fvm_get_csr_size(fvm_handle h, int64_t* n_rows, int64_t* nnz) {
    NEED_ASSEMBLED(h);
    if (n_rows) *n_rows = h->csr.n;
    if (nnz) *nnz = h->csr.nnz;
    return FVM_OK;
}

extern "C" int32_t fvm_get_csr(fvm_handle h, int32_t* rowptr, int32_t* col, double* val, double* b) {
    NEED_ASSEMBLED(h);
    FVM_REQUIRE(h, rowptr && col && val, "fvm_get_csr: null argument");
    const Csr& c = h->csr;
    const int64_t N = c.n, nnz = c.nnz;
    std::vector<int32_t> rp(N + 1), cl(nnz);
    std::vector<double> vl(nnz), bb(N);
    FVM_CUDA(h, cudaMemcpy(rp.data(), c.rowptr, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost));
    FVM_CUDA(h, cudaMemcpy(cl.data(), c.col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost));
    FVM_CUDA(h, cudaMemcpy(vl.data(), c.val, sizeof(double) * nnz, cudaMemcpyDeviceToHost));
    FVM_CUDA(h, cudaMemcpy(bb.data(), c.b, sizeof(double) * N, cudaMemcpyDeviceToHost));
    const int32_t* old_of_new = h->node_old_of_new.data();
    const int32_t* new_of_old = h->node_new_of_old.data();
    rowptr[0] = 0;
    for (int64_t o = 0; o < N; ++o) {
        const int32_t g = new_of_old[o];
        rowptr[o + 1] = rowptr[o] + (rp[g + 1] - rp[g]);
    }
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < N; ++o) {
        const int32_t g = new_of_old[o];
        const int len = rp[g + 1] - rp[g];
        int32_t* oc = col + rowptr[o];
        double* ov = val + rowptr[o];
        for (int k = 0; k < len; ++k) {  // insertion sort by caller column id
            const int32_t cc = old_of_new[cl[rp[g] + k]];
            const double vv = vl[rp[g] + k];
            int pos = k;
            while (pos > 0 && oc[pos - 1] > cc) {
                oc[pos] = oc[pos - 1];
                ov[pos] = ov[pos - 1];
                --pos;
            }
            oc[pos] = cc;
            ov[pos] = vv;
        }
        if (b) b[o] = bb[g];
    }
    return FVM_OK;
}

extern "C" int32_t fvm_spmv_native(fvm_handle h, const double* x, double* y, int32_t add_b) {
    NEED_ASSEMBLED(h);
    FVM_REQUIRE(h, x && y && x != y, "fvm_spmv_native: bad arguments");
    return fvm_apply_spmv(h, const_cast<double*>(x), y, add_b != 0, false);
}

extern "C" int32_t fvm_spmv(fvm_handle h, const double* x, double* y, int32_t add_b, int32_t on_device) {
    NEED_ASSEMBLED(h);
    FVM_REQUIRE(h, x && y, "fvm_spmv: null argument");
    int32_t rc = fvm_ensure_state(h);
    if (rc) return rc;
    const auto wall0 = std::chrono::steady_clock::now();
    if (!on_device && x != y) {  // large host vectors: the banded copy / compute pipeline of fvm_pipe.cu
        bool used = false;
        rc = fvm_spmv_pipelined(h, x, y, add_b != 0, &used);
        if (rc || used) {
            if (!rc) fvm_pipe_report(h, 1, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
            return rc;
        }
    }
    const size_t bytes = sizeof(double) * h->N;
    const double* src = x;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, x, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, h->d_u, true))) return rc;
    if ((rc = fvm_apply_spmv(h, h->d_u, h->d_du, add_b != 0, false))) return rc;
    if (on_device) {
        if ((rc = fvm_launch_permute(h, h->d_du, y, false))) return rc;
    } else {
        if ((rc = fvm_launch_permute(h, h->d_du, h->d_io, false))) return rc;
        FVM_CUDA(h, cudaMemcpyAsync(y, h->d_io, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!on_device && x != y) fvm_pipe_report(h, 1, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
    return FVM_OK;
}
