// libfvmcuda: the sparse Jacobian d(fvm_eqs!)/du on the device.
//
// Pattern = jacobian_sparsity (/root/reference/src/solve.jl:56-77 scalar, :96-131 system: node-major
// interleaving (i-1)*neq + l, dense neq x neq blocks).  Values are exact derivatives: the flux / source /
// condition registry is evaluated on forward-mode dual numbers (the reference pushes ForwardDiff duals
// through fvm_eqs!, src/solve.jl:5; Julia closures cannot run here, the registry can).  One thread owns a
// row block (node) and gathers its incident triangles and boundary edges in a fixed order: no atomics.
#include <algorithm>
#include <cstring>

#include "fvm_device.cuh"

#define JAC_MAX_ROW 64

struct JacEdge {  // one boundary edge (native ids), for the rows of its two endpoints
    int32_t v[3];
    int32_t i, j;
    uint8_t kind[FVM_MAX_NEQ];
    int32_t fidx[FVM_MAX_NEQ];
    double Di, Dj;  // tabulated D at the quarter points (FVM_FLUX_DIFF_TABLE)
};

template <int NEQ, class T>
__device__ __forceinline__ void flux_dispatch_t(const FluxParams& fp, double x, double y, double t, const T* a, const T* b,
                                                const T* g, double dtab, T* qx, T* qy) {
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

struct JacArgs {
    const int32_t* n2t_ptr;
    const int32_t* n2t;
    const int32_t* tri;
    const int32_t* rowptr;
    const int32_t* col;
    const double* dtab_native;  // [3][tpad] (DevMesh::dtab) or null
    double* val;                // [nnz][NEQ][NEQ]
    // boundary edges by node
    const JacEdge* edges;
    const int32_t* bn_of_node;  // [N] index into bn_ptr or -1
    const int32_t* bn_ptr;
    const int32_t* bn_items;    // edge << 1 | role
};

template <int NEQ>
__global__ void __launch_bounds__(128)
    jacobian_rows_kernel(const DevMesh m, const FluxParams fp, const SourceParams sp, const JacArgs a, const double t,
                         const double* __restrict__ u) {
    using D3 = Dual<3 * NEQ>;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m.n_nodes) return;
    const int beg = a.rowptr[g], len = a.rowptr[g + 1] - beg;
    double* __restrict__ row = a.val + (size_t)beg * NEQ * NEQ;
    for (int k = 0; k < len * NEQ * NEQ; ++k) row[k] = 0.0;
    if (g >= m.n_vertices) return;  // not a vertex: du = 0
    uint8_t kind[NEQ];
    bool any_free = false;
#pragma unroll
    for (int l = 0; l < NEQ; ++l) {
        kind[l] = m.kind[(size_t)l * m.n_nodes + g];
        any_free = any_free || kind[l] == FVM_NODE_FREE;
    }
    const double invV = 1.0 / m.vol[g];
    int dpos = 0;
    for (int k = 0; k < len; ++k)
        if (a.col[beg + k] == g) dpos = k;
    auto add_block = [&](int colnode, int l, int lp, double v) {
        int pos = 0;
        for (int k = 0; k < len; ++k)
            if (a.col[beg + k] == colnode) pos = k;
        row[(pos * NEQ + l) * NEQ + lp] += v;
    };
    if (any_free) {
        // ---- triangles, triangle_contributions.jl:28-35 ------------------------------------------
        for (int e = a.n2t_ptr[g]; e < a.n2t_ptr[g + 1]; ++e) {
            const int code = a.n2t[e];
            const int64_t tr = code >> 2;
            const int p = code & 3;
            const int v[3] = {a.tri[3 * tr], a.tri[3 * tr + 1], a.tri[3 * tr + 2]};
            TriGeom G;
            tri_geometry<false>(m.xy[2 * (size_t)v[0]], m.xy[2 * (size_t)v[0] + 1], m.xy[2 * (size_t)v[1]], m.xy[2 * (size_t)v[1] + 1],
                                m.xy[2 * (size_t)v[2]], m.xy[2 * (size_t)v[2] + 1], G, nullptr);
            D3 al[NEQ], be[NEQ], ga[NEQ];
#pragma unroll
            for (int l = 0; l < NEQ; ++l) {
                al[l] = D3(0.0);
                be[l] = D3(0.0);
                ga[l] = D3(0.0);
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    const double um = u[(size_t)v[mm] * NEQ + l];
                    al[l].v += G.s[mm] * um;
                    be[l].v += G.s[3 + mm] * um;
                    ga[l].v += G.s[6 + mm] * um;
                    al[l].d[mm * NEQ + l] = G.s[mm];
                    be[l].d[mm * NEQ + l] = G.s[3 + mm];
                    ga[l].d[mm * NEQ + l] = G.s[6 + mm];
                }
            }
            // contribution to this node: Q_{(p+2)%3} - Q_p
            D3 c[NEQ];
#pragma unroll
            for (int l = 0; l < NEQ; ++l) c[l] = D3(0.0);
            for (int s = 0; s < 2; ++s) {
                const int ed = s == 0 ? (p + 2) % 3 : p;
                const double sign = s == 0 ? 1.0 : -1.0;
                D3 qx[NEQ], qy[NEQ];
                const double dt = a.dtab_native ? a.dtab_native[(size_t)ed * m.tpad + tr] : 0.0;
                flux_dispatch_t<NEQ, D3>(fp, G.mx[ed], G.my[ed], t, al, be, ga, dt, qx, qy);
#pragma unroll
                for (int l = 0; l < NEQ; ++l) c[l] = c[l] + (qx[l] * G.ey[ed] + qy[l] * (-G.ex[ed])) * sign;
            }
#pragma unroll
            for (int l = 0; l < NEQ; ++l) {
                if (kind[l] != FVM_NODE_FREE) continue;
                for (int mm = 0; mm < 3; ++mm)
#pragma unroll
                    for (int lp = 0; lp < NEQ; ++lp) add_block(v[mm], l, lp, c[l].d[mm * NEQ + lp] * invV);
            }
        }
        // ---- boundary edges, boundary_edge_contributions.jl:41-86: du[i] -= Q ---------------------
        const int bq = a.bn_of_node ? a.bn_of_node[g] : -1;
        if (bq >= 0) {
            for (int it = a.bn_ptr[bq]; it < a.bn_ptr[bq + 1]; ++it) {
                const JacEdge E = a.edges[a.bn_items[it] >> 1];
                const int role = a.bn_items[it] & 1;
                TriGeom G;
                tri_geometry<false>(m.xy[2 * (size_t)E.v[0]], m.xy[2 * (size_t)E.v[0] + 1], m.xy[2 * (size_t)E.v[1]],
                                    m.xy[2 * (size_t)E.v[1] + 1], m.xy[2 * (size_t)E.v[2]], m.xy[2 * (size_t)E.v[2] + 1], G, nullptr);
                D3 al[NEQ], be[NEQ], ga[NEQ];
#pragma unroll
                for (int l = 0; l < NEQ; ++l) {
                    al[l] = D3(0.0);
                    be[l] = D3(0.0);
                    ga[l] = D3(0.0);
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) {
                        const double um = u[(size_t)E.v[mm] * NEQ + l];
                        al[l].v += G.s[mm] * um;
                        be[l].v += G.s[3 + mm] * um;
                        ga[l].v += G.s[6 + mm] * um;
                        al[l].d[mm * NEQ + l] = G.s[mm];
                        be[l].d[mm * NEQ + l] = G.s[3 + mm];
                        ga[l].d[mm * NEQ + l] = G.s[6 + mm];
                    }
                }
                const double px = m.xy[2 * (size_t)E.i], py = m.xy[2 * (size_t)E.i + 1];
                const double qxx = m.xy[2 * (size_t)E.j], qyy = m.xy[2 * (size_t)E.j + 1];
                const double dx = qxx - px, dy = qyy - py;
                const double lij = sqrt(dx * dx + dy * dy);
                const double nx = dy / lij, ny = -dx / lij, lh = 0.5 * lij;
                const double mijx = 0.5 * (px + qxx), mijy = 0.5 * (py + qyy);
                const double X = role ? 0.5 * (qxx + mijx) : 0.5 * (px + mijx);
                const double Y = role ? 0.5 * (qyy + mijy) : 0.5 * (py + mijy);
                D3 qx[NEQ], qy[NEQ];
                flux_dispatch_t<NEQ, D3>(fp, X, Y, t, al, be, ga, role ? E.Dj : E.Di, qx, qy);
#pragma unroll
                for (int l = 0; l < NEQ; ++l) {
                    if (kind[l] != FVM_NODE_FREE) continue;
                    D3 Q;
                    if (E.kind[l] == FVM_EDGE_NEUMANN) {
                        const CondFn cf = m.cond[l * FVM_MAX_COND_FN + E.fidx[l]];
                        const D3 ushape = al[l] * X + be[l] * Y + ga[l];
                        Q = cond_eval_t<D3>(cf, X, Y, t, ushape) * lh;
                    } else {
                        Q = (qx[l] * nx + qy[l] * ny) * lh;
                    }
                    for (int mm = 0; mm < 3; ++mm)
#pragma unroll
                        for (int lp = 0; lp < NEQ; ++lp) add_block(E.v[mm], l, lp, -Q.d[mm * NEQ + lp] * invV);
                }
            }
        }
    }
    // ---- node pass, source_contributions.jl:33-68 -------------------------------------------------
    using DN = Dual<NEQ>;
    DN un[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) {
        un[l] = DN(u[(size_t)g * NEQ + l]);
        un[l].d[l] = 1.0;
    }
    double tab[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) tab[l] = (sp.model == FVM_SRC_TABLE && m.src_tab) ? m.src_tab[(size_t)g * NEQ + l] : 0.0;
#pragma unroll
    for (int l = 0; l < NEQ; ++l) {
        if (kind[l] == FVM_NODE_FREE) {
            const DN S = source_eval_t<NEQ, DN>(sp, l, un, tab);
#pragma unroll
            for (int lp = 0; lp < NEQ; ++lp) row[(dpos * NEQ + l) * NEQ + lp] += S.d[lp];
        } else if (kind[l] == FVM_NODE_DUDT) {
            const CondFn cf = m.cond[l * FVM_MAX_COND_FN + m.fidx[(size_t)l * m.n_nodes + g]];
            const DN r = cond_eval_t<DN>(cf, m.xy[2 * (size_t)g], m.xy[2 * (size_t)g + 1], t, un[l]);
#pragma unroll
            for (int lp = 0; lp < NEQ; ++lp) row[(dpos * NEQ + l) * NEQ + lp] = r.d[lp];
        }
    }
}

int32_t fvm_build_pattern(fvm_ctx* h);  // fvm_linear.cu

#define NEED_FINAL(h)                                                                       \
    do {                                                                                    \
        if (!(h)) return FVM_ERR_ARG;                                                       \
        if (!(h)->finalized) return fvm_fail((h), FVM_ERR_STATE, "call fvm_finalize first"); \
        FVM_CUDA(h, cudaSetDevice((h)->device));                                            \
    } while (0)

static int32_t jac_prepare(fvm_ctx* h) {
    if (h->jac_val) return FVM_OK;
    int32_t rc = fvm_build_pattern(h);
    if (rc) return rc;
    const int64_t N = h->N, Eb = h->Eb;
    const int neq = h->neq;
    if ((rc = fvm_dev_alloc(h, &h->jac_val, (size_t)h->csr.nnz * neq * neq))) return rc;
    if (Eb > 0) {
        std::vector<JacEdge> edges(Eb);
        std::vector<std::pair<int32_t, int32_t>> items;
        for (int64_t e = 0; e < Eb; ++e) {
            JacEdge& E = edges[e];
            const int32_t* v = h->h_tri.data() + 3 * (int64_t)h->h_edge_tri[e];
            for (int q = 0; q < 3; ++q) E.v[q] = h->node_new_of_old[v[q]];
            E.i = h->node_new_of_old[h->h_bedge[2 * e]];
            E.j = h->node_new_of_old[h->h_bedge[2 * e + 1]];
            for (int l = 0; l < FVM_MAX_NEQ; ++l) {
                E.kind[l] = l < neq ? h->h_ekind[l][e] : 0;
                E.fidx[l] = l < neq ? h->h_efidx[l][e] : 0;
            }
            E.Di = h->h_dbnd.empty() ? 0.0 : h->h_dbnd[2 * e];
            E.Dj = h->h_dbnd.empty() ? 0.0 : h->h_dbnd[2 * e + 1];
            items.push_back({E.i, (int32_t)(e << 1)});
            items.push_back({E.j, (int32_t)(e << 1 | 1)});
        }
        std::stable_sort(items.begin(), items.end(), [](auto& x, auto& y) { return x.first < y.first; });
        std::vector<int32_t> bn_of(N, -1), bn_ptr, bn_items;
        for (size_t k = 0; k < items.size(); ++k) {
            if (k == 0 || items[k].first != items[k - 1].first) {
                bn_of[items[k].first] = (int32_t)bn_ptr.size();
                bn_ptr.push_back((int32_t)k);
            }
            bn_items.push_back(items[k].second);
        }
        bn_ptr.push_back((int32_t)items.size());
        JacEdge* d_e = nullptr;
        if ((rc = fvm_dev_upload(h, &d_e, edges))) return rc;
        h->jac_edges = d_e;
        if ((rc = fvm_dev_upload(h, &h->jac_bn_of, bn_of))) return rc;
        if ((rc = fvm_dev_upload(h, &h->jac_bn_ptr, bn_ptr))) return rc;
        if ((rc = fvm_dev_upload(h, &h->jac_bn_items, bn_items))) return rc;
    }
    return FVM_OK;
}

// block values of d(fvm_eqs!)/du at the native-order state `u_native` (ghost layer already refreshed) -> h->jac_val
int32_t fvm_launch_jacobian(fvm_ctx* h, double t, const double* u_native) {
    if (h->neq > 2) return fvm_fail(h, FVM_ERR_ARG, "fvm_jacobian: systems with more than 2 species are not compiled");
    int32_t rc = jac_prepare(h);
    if (rc) return rc;
    JacArgs a{};
    a.n2t_ptr = h->csr.n2t_ptr;
    a.n2t = h->csr.n2t;
    a.tri = h->d_tri_native;
    a.rowptr = h->csr.rowptr;
    a.col = h->csr.col;
    a.dtab_native = h->dm.dtab;
    a.val = h->jac_val;
    a.edges = (const JacEdge*)h->jac_edges;
    a.bn_of_node = h->jac_bn_of;
    a.bn_ptr = h->jac_bn_ptr;
    a.bn_items = h->jac_bn_items;
    const unsigned grid = (unsigned)((h->N + 127) / 128);
    if (h->neq == 1) jacobian_rows_kernel<1><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, h->source, a, t, u_native);
    else jacobian_rows_kernel<2><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, h->source, a, t, u_native);
    FVM_CUDA(h, cudaGetLastError());
    h->jac_ready = true;
    return FVM_OK;
}

extern "C" int32_t fvm_jacobian(fvm_handle h, double t, const double* u, int32_t on_device) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, u, "fvm_jacobian: null state");
    FVM_REQUIRE(h, h->neq <= 2, "fvm_jacobian: systems with more than 2 species are not compiled");
    int32_t rc = jac_prepare(h);
    if (rc) return rc;
    if ((rc = fvm_ensure_state(h))) return rc;
    const size_t bytes = sizeof(double) * h->N * h->neq;
    const double* src = u;
    if (!on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, h->d_u, true))) return rc;
    if ((rc = fvm_halo_exchange(h, h->d_u))) return rc;
    if ((rc = fvm_launch_jacobian(h, t, h->d_u))) return rc;
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FVM_OK;
}

extern "C" int32_t fvm_get_jacobian_size(fvm_handle h, int64_t* n_rows, int64_t* nnz) {
    NEED_FINAL(h);
    int32_t rc = jac_prepare(h);
    if (rc) return rc;
    if (n_rows) *n_rows = h->N * h->neq;
    if (nnz) *nnz = h->csr.nnz * h->neq * h->neq;
    return FVM_OK;
}

// caller numbering, node-major interleaving (row = i*neq + l, col = j*neq + l'), columns sorted
extern "C" int32_t fvm_get_jacobian_csr(fvm_handle h, int32_t* rowptr, int32_t* col, double* val) {
    NEED_FINAL(h);
    FVM_REQUIRE(h, rowptr && col, "fvm_get_jacobian_csr: null argument");
    if (val && !h->jac_ready) return fvm_fail(h, FVM_ERR_STATE, "fvm_get_jacobian_csr: call fvm_jacobian first");
    int32_t rc = jac_prepare(h);
    if (rc) return rc;
    const int64_t N = h->N, nnz = h->csr.nnz;
    const int neq = h->neq;
    std::vector<int32_t> rp(N + 1), cl(nnz);
    std::vector<double> vl;
    FVM_CUDA(h, cudaMemcpy(rp.data(), h->csr.rowptr, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost));
    FVM_CUDA(h, cudaMemcpy(cl.data(), h->csr.col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost));
    if (val) {
        vl.resize((size_t)nnz * neq * neq);
        FVM_CUDA(h, cudaMemcpy(vl.data(), h->jac_val, sizeof(double) * vl.size(), cudaMemcpyDeviceToHost));
    }
    const int32_t* old_of_new = h->node_old_of_new.data();
    const int32_t* new_of_old = h->node_new_of_old.data();
    rowptr[0] = 0;
    for (int64_t o = 0; o < N; ++o) {
        const int32_t g = new_of_old[o];
        const int len = rp[g + 1] - rp[g];
        for (int l = 0; l < neq; ++l) rowptr[o * neq + l + 1] = rowptr[o * neq + l] + len * neq;
    }
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < N; ++o) {
        const int32_t g = new_of_old[o];
        const int len = rp[g + 1] - rp[g];
        int32_t order[JAC_MAX_ROW];
        for (int k = 0; k < len; ++k) {  // sort the block columns by caller node id
            const int32_t cc = old_of_new[cl[rp[g] + k]];
            int pos = k;
            while (pos > 0 && old_of_new[cl[rp[g] + order[pos - 1]]] > cc) {
                order[pos] = order[pos - 1];
                --pos;
            }
            order[pos] = k;
        }
        for (int l = 0; l < neq; ++l) {
            int32_t* oc = col + rowptr[o * neq + l];
            double* ov = val ? val + rowptr[o * neq + l] : nullptr;
            for (int k = 0; k < len; ++k) {
                const int src = order[k];
                const int32_t cc = old_of_new[cl[rp[g] + src]];
                for (int lp = 0; lp < neq; ++lp) {
                    oc[k * neq + lp] = cc * neq + lp;
                    if (ov) ov[k * neq + lp] = vl[((size_t)(rp[g] + src) * neq + l) * neq + lp];
                }
            }
        }
    }
    return FVM_OK;
}
