// libfvmcuda: post-processing on the device: piecewise-linear interpolation
// (/root/reference/src/utils.jl:23-27) and edge fluxes (compute_flux, src/problem.jl:458-487).
#include "fvm_device.cuh"

template <int NEQ>
__global__ void post_kernel(const DevMesh m, const FluxParams fp, const double t, const int32_t* __restrict__ tri_native,
                            const double* __restrict__ u, const int64_t n, const int32_t* __restrict__ tri_q,
                            const double* __restrict__ xy_q, const double* __restrict__ nrm_q, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t tr = tri_q[k];
    const int v[3] = {tri_native[3 * tr], tri_native[3 * tr + 1], tri_native[3 * tr + 2]};
    TriGeom G;
    tri_geometry<true>(m.xy[2 * (size_t)v[0]], m.xy[2 * (size_t)v[0] + 1], m.xy[2 * (size_t)v[1]], m.xy[2 * (size_t)v[1] + 1],
                       m.xy[2 * (size_t)v[2]], m.xy[2 * (size_t)v[2] + 1], G, nullptr);
    double a[NEQ], b[NEQ], g[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) {
        const double ui = u[(size_t)v[0] * NEQ + l], uj = u[(size_t)v[1] * NEQ + l], uk = u[(size_t)v[2] * NEQ + l];
        a[l] = G.s[0] * ui + G.s[1] * uj + G.s[2] * uk;  // shape_functions.jl:2-19
        b[l] = G.s[3] * ui + G.s[4] * uj + G.s[5] * uk;
        g[l] = G.s[6] * ui + G.s[7] * uj + G.s[8] * uk;
    }
    const double x = xy_q[2 * k], y = xy_q[2 * k + 1];
    if (!nrm_q) {  // pl_interpolate: alpha x + beta y + gamma
#pragma unroll
        for (int l = 0; l < NEQ; ++l) out[k * NEQ + l] = a[l] * x + b[l] * y + g[l];
        return;
    }
    double qx[NEQ], qy[NEQ];
    switch (fp.model) {
        case FVM_FLUX_DIFF_POWER: flux_eval<FVM_FLUX_DIFF_POWER, NEQ>(fp, x, y, t, a, b, g, 0.0, qx, qy); break;
        case FVM_FLUX_ADVDIFF: flux_eval<FVM_FLUX_ADVDIFF, NEQ>(fp, x, y, t, a, b, g, 0.0, qx, qy); break;
        case FVM_FLUX_KELLER_SEGEL:
            if constexpr (NEQ == 2) flux_eval<FVM_FLUX_KELLER_SEGEL, NEQ>(fp, x, y, t, a, b, g, 0.0, qx, qy);
            break;
        default: flux_eval<FVM_FLUX_DIFF_CONST, NEQ>(fp, x, y, t, a, b, g, 0.0, qx, qy); break;
    }
#pragma unroll
    for (int l = 0; l < NEQ; ++l) out[k * NEQ + l] = nrm_q[2 * k] * qx[l] + nrm_q[2 * k + 1] * qy[l];
}

// Evaluates, for n query points lying in given triangles (caller triangle indices): the piecewise-linear
// interpolant of u (nrm == NULL), or the flux q(x, y, t, alpha, beta, gamma) . nrm.  Query arrays and
// `out` are host buffers; u is host or device (caller order).
extern "C" int32_t fvm_eval_points(fvm_handle h, double t, const double* u, int32_t u_on_device, int64_t n,
                                   const int32_t* tri_idx, const double* xy, const double* nrm, double* out) {
    if (!h) return FVM_ERR_ARG;
    if (!h->finalized) return fvm_fail(h, FVM_ERR_STATE, "call fvm_finalize first");
    FVM_CUDA(h, cudaSetDevice(h->device));
    FVM_REQUIRE(h, u && n >= 0 && (n == 0 || (tri_idx && xy && out)), "fvm_eval_points: null argument");
    if (nrm && h->flux.model == FVM_FLUX_DIFF_TABLE)
        return fvm_fail(h, FVM_ERR_UNSUPPORTED, "fvm_eval_points: a tabulated D(x,y) is only known at the cv-edge midpoints");
    if (n == 0) return FVM_OK;
    int32_t rc = fvm_ensure_state(h);
    if (rc) return rc;
    if (h->tri_new_of_old.empty()) {
        h->tri_new_of_old.resize(h->T);
        for (int64_t nt = 0; nt < h->T; ++nt) h->tri_new_of_old[h->tri_old_of_new[nt]] = (int32_t)nt;
    }
    std::vector<int32_t> tq(n);
    for (int64_t k = 0; k < n; ++k) {
        const int64_t o = (int64_t)tri_idx[k] - h->h_index_base;
        FVM_REQUIRE(h, o >= 0 && o < h->T, "fvm_eval_points: triangle index out of range");
        tq[k] = h->tri_new_of_old[o];
    }
    const size_t bytes = sizeof(double) * h->N * h->neq;
    const double* src = u;
    if (!u_on_device) {
        FVM_CUDA(h, cudaMemcpyAsync(h->d_io, u, bytes, cudaMemcpyHostToDevice, h->stream));
        src = h->d_io;
    }
    if ((rc = fvm_launch_permute(h, src, h->d_u, true))) return rc;
    int32_t* d_t = nullptr;
    double *d_xy = nullptr, *d_n = nullptr, *d_o = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_t, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_xy, sizeof(double) * 2 * n);
    if (e == cudaSuccess && nrm) e = cudaMalloc((void**)&d_n, sizeof(double) * 2 * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, sizeof(double) * n * h->neq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, tq.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xy, xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && nrm) e = cudaMemcpyAsync(d_n, nrm, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)((n + 127) / 128);
        switch (h->neq) {
            case 1: post_kernel<1><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, t, h->d_tri_native, h->d_u, n, d_t, d_xy, d_n, d_o); break;
            case 2: post_kernel<2><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, t, h->d_tri_native, h->d_u, n, d_t, d_xy, d_n, d_o); break;
            case 3: post_kernel<3><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, t, h->d_tri_native, h->d_u, n, d_t, d_xy, d_n, d_o); break;
            default: post_kernel<4><<<grid, 128, 0, h->stream>>>(h->dm, h->flux, t, h->d_tri_native, h->d_u, n, d_t, d_xy, d_n, d_o); break;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, sizeof(double) * n * h->neq, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_t);
    cudaFree(d_xy);
    cudaFree(d_n);
    cudaFree(d_o);
    FVM_CUDA(h, e);
    return FVM_OK;
}
