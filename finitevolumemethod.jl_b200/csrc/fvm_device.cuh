// Device-side building blocks shared by the RHS, geometry and assembly kernels.
#pragma once
#include "fvm_internal.h"

// ---- arithmetic policy: EXACT = every operation individually rounded (no FMA contraction),
// so the geometry is bit-identical to the reference's Julia arithmetic; FAST lets nvcc contract.
template <bool EXACT>
struct Ar {
    static __device__ __forceinline__ double mul(double a, double b) {
        if constexpr (EXACT) return __dmul_rn(a, b);
        else return a * b;
    }
    static __device__ __forceinline__ double add(double a, double b) {
        if constexpr (EXACT) return __dadd_rn(a, b);
        else return a + b;
    }
    static __device__ __forceinline__ double sub(double a, double b) {
        if constexpr (EXACT) return __dsub_rn(a, b);
        else return a - b;
    }
    static __device__ __forceinline__ double div(double a, double b) {
        if constexpr (EXACT) return __ddiv_rn(a, b);
        else return a / b;
    }
};

// x / 3 correctly rounded without the IEEE division sequence (Markstein: q = RN(x/3) from two FMAs on a
// faithful first guess; exact for every double outside the subnormal range, brute-forced on 5e8 inputs):
// the centroid (geometry.jl:114) keeps the reference's rounding, which matters because the cv-edge vectors
// c - m_e cancel to h/6 and would otherwise differ at 1e-12
__device__ __forceinline__ double div3_rn(double x) {
    const double y = 1.0 / 3.0;
    const double q = x * y;
    const double r = fma(-3.0, q, x);
    return fma(r, y, q);
}
// 1/x to ~1 ulp for any normal double: MUFU.RCP64H seed + two Newton steps (no slow-path branch)
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

struct TriGeom {
    double s[9];          // shape-function coefficients s1..s9
    double mx[3], my[3];  // cv-edge midpoints
    double ex[3], ey[3];  // centroid - edge midpoint; length-scaled normal is (ey, -ex)
};

// /root/reference/src/geometry.jl:107-161, same operation order.  EXACT: every operation individually rounded (IEEE
// divisions).  !EXACT: contracted FMAs where the result is provably the same, a Newton reciprocal of Delta, and
// EXACT_S selects how s1..s6 are formed from it (s7..s9 are always the correctly rounded quotients).
template <bool EXACT, bool EXACT_S = EXACT>
__device__ __forceinline__ void tri_geometry(double px, double py, double qx, double qy, double rx, double ry,
                                             TriGeom& G, double* S /* 3 sub-cv areas or nullptr */) {
    using A = Ar<EXACT>;
    double cx, cy;
    if constexpr (EXACT) {
        cx = A::div(A::add(A::add(px, qx), rx), 3.0);
        cy = A::div(A::add(A::add(py, qy), ry), 3.0);
    } else {  // same value, 3 instructions instead of the division sequence
        cx = div3_rn(__dadd_rn(__dadd_rn(px, qx), rx));
        cy = div3_rn(__dadd_rn(__dadd_rn(py, qy), ry));
    }
    const double m1x = A::mul(A::add(px, qx), 0.5), m1y = A::mul(A::add(py, qy), 0.5);
    const double m2x = A::mul(A::add(qx, rx), 0.5), m2y = A::mul(A::add(qy, ry), 0.5);
    const double m3x = A::mul(A::add(rx, px), 0.5), m3y = A::mul(A::add(ry, py), 0.5);
    if (S) {
        const double pcx = A::sub(cx, px), pcy = A::sub(cy, py);
        const double qcx = A::sub(cx, qx), qcy = A::sub(cy, qy);
        const double rcx = A::sub(cx, rx), rcy = A::sub(cy, ry);
        const double m13x = A::sub(m1x, m3x), m13y = A::sub(m1y, m3y);
        const double m21x = A::sub(m2x, m1x), m21y = A::sub(m2y, m1y);
        const double m32x = A::sub(m3x, m2x), m32y = A::sub(m3y, m2y);
        S[0] = 0.5 * fabs(A::sub(A::mul(pcx, m13y), A::mul(pcy, m13x)));
        S[1] = 0.5 * fabs(A::sub(A::mul(qcx, m21y), A::mul(qcy, m21x)));
        S[2] = 0.5 * fabs(A::sub(A::mul(rcx, m32y), A::mul(rcy, m32x)));
    }
    // Delta = qx*ry - qy*rx - px*ry + rx*py + px*qy - qx*py   (left to right).  Delta and the numerators
    // of s7..s9 cancel catastrophically in absolute coordinates (relative error ~ eps * |x|^2 / area:
    // 1e-11 at h = 2e-3, 1e-9 on the 4096^2 lattice), so they are ALWAYS evaluated with the reference's
    // individually rounded operations -- a contracted FMA would change the result at that level.
    using E = Ar<true>;
    double D = E::sub(E::mul(qx, ry), E::mul(qy, rx));
    D = E::sub(D, E::mul(px, ry));
    D = E::add(D, E::mul(rx, py));
    D = E::add(D, E::mul(px, qy));
    D = E::sub(D, E::mul(qx, py));
    const double n7 = E::sub(E::mul(qx, ry), E::mul(rx, qy));
    const double n8 = E::sub(E::mul(rx, py), E::mul(px, ry));
    const double n9 = E::sub(E::mul(px, qy), E::mul(qx, py));
    if constexpr (EXACT) {
        G.s[0] = A::div(A::sub(qy, ry), D);
        G.s[1] = A::div(A::sub(ry, py), D);
        G.s[2] = A::div(A::sub(py, qy), D);
        G.s[3] = A::div(A::sub(rx, qx), D);
        G.s[4] = A::div(A::sub(px, rx), D);
        G.s[5] = A::div(A::sub(qx, px), D);
        G.s[6] = A::div(n7, D);
        G.s[7] = A::div(n8, D);
        G.s[8] = A::div(n9, D);
    } else {
        // s1..s6 = num * RN(1/D): within an ulp of the reference's quotient, harmless (6.5e-16 on du at 4096^2).
        // s7..s9 ~ |x| / h are different: gamma = s7 u_i + s8 u_j + s9 u_k cancels from ~|u| |x| / h down to |u|, so an ulp
        // there is 2e-12 in a u-dependent flux on the 4096^2 lattice.  They get the correctly rounded quotient without
        // the IEEE division sequence (Markstein): with r = RN(1/D) and q0 = RN(n r), RN(q0 + r RN(n - D q0)) = RN(n / D)
        const double iD = rcp_newton(D);
        auto quot = [&](double n) {
            const double q0 = n * iD;
            return fma(fma(-D, q0, n), iD, q0);
        };
        if constexpr (EXACT_S) {  // u-dependent fluxes: alpha x + beta y + gamma cancels by |x| / h, every ulp of s counts
            G.s[0] = quot(qy - ry);
            G.s[1] = quot(ry - py);
            G.s[2] = quot(py - qy);
            G.s[3] = quot(rx - qx);
            G.s[4] = quot(px - rx);
            G.s[5] = quot(qx - px);
        } else {
            G.s[0] = (qy - ry) * iD;
            G.s[1] = (ry - py) * iD;
            G.s[2] = (py - qy) * iD;
            G.s[3] = (rx - qx) * iD;
            G.s[4] = (px - rx) * iD;
            G.s[5] = (qx - px) * iD;
        }
        G.s[6] = quot(n7);
        G.s[7] = quot(n8);
        G.s[8] = quot(n9);
    }
    G.mx[0] = A::mul(A::add(m1x, cx), 0.5);
    G.my[0] = A::mul(A::add(m1y, cy), 0.5);
    G.mx[1] = A::mul(A::add(m2x, cx), 0.5);
    G.my[1] = A::mul(A::add(m2y, cy), 0.5);
    G.mx[2] = A::mul(A::add(m3x, cx), 0.5);
    G.my[2] = A::mul(A::add(m3y, cy), 0.5);
    G.ex[0] = A::sub(cx, m1x);
    G.ey[0] = A::sub(cy, m1y);
    G.ex[1] = A::sub(cx, m2x);
    G.ey[1] = A::sub(cy, m2y);
    G.ex[2] = A::sub(cx, m3x);
    G.ey[2] = A::sub(cy, m3y);
}

// 1/x to fp64 round-off without the ~20-instruction IEEE division sequence: fp32 seed + two Newton steps
// (relative error < 2^-51; the parity tolerance is 1e-12)
__device__ __forceinline__ double fast_rcp(double x) {
    double r = (double)__frcp_rn((float)x);
    r = r * (2.0 - x * r);
    r = r * (2.0 - x * r);
    return r;
}

// ---- forward-mode dual numbers: the analytic Jacobian of fvm_eqs! reuses the flux / source /
// condition registry below unchanged
template <int K>
struct Dual {
    double v;
    double d[K];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double c) : v(c) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = 0.0;
    }
};
#define DUAL_BIN(op, vexpr, dexpr)                                                            \
    template <int K>                                                                          \
    __device__ __forceinline__ Dual<K> operator op(const Dual<K>& a, const Dual<K>& b) {      \
        Dual<K> r;                                                                            \
        r.v = vexpr;                                                                          \
        _Pragma("unroll") for (int k = 0; k < K; ++k) r.d[k] = dexpr;                         \
        return r;                                                                             \
    }
DUAL_BIN(+, a.v + b.v, a.d[k] + b.d[k])
DUAL_BIN(-, a.v - b.v, a.d[k] - b.d[k])
DUAL_BIN(*, a.v * b.v, a.d[k] * b.v + a.v * b.d[k])
#undef DUAL_BIN
template <int K> __device__ __forceinline__ Dual<K> operator+(const Dual<K>& a, double b) { Dual<K> r = a; r.v += b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator+(double b, const Dual<K>& a) { return a + b; }
template <int K> __device__ __forceinline__ Dual<K> operator-(const Dual<K>& a, double b) { Dual<K> r = a; r.v -= b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator-(const Dual<K>& a) {
    Dual<K> r;
    r.v = -a.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
    return r;
}
template <int K> __device__ __forceinline__ Dual<K> operator-(double b, const Dual<K>& a) { return (-a) + b; }
template <int K> __device__ __forceinline__ Dual<K> operator*(const Dual<K>& a, double b) {
    Dual<K> r;
    r.v = a.v * b;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b;
    return r;
}
template <int K> __device__ __forceinline__ Dual<K> operator*(double b, const Dual<K>& a) { return a * b; }
template <int K> __device__ __forceinline__ Dual<K> operator/(const Dual<K>& a, double b) { return a * (1.0 / b); }

__device__ __forceinline__ double fabs_t(double x) { return fabs(x); }
__device__ __forceinline__ double pow_t(double x, double e) { return pow(x, e); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
template <int K> __device__ __forceinline__ Dual<K> fabs_t(const Dual<K>& a) { return a.v < 0.0 ? -a : a; }
template <int K> __device__ __forceinline__ Dual<K> pow_t(const Dual<K>& a, double e) {
    Dual<K> r;
    r.v = pow(a.v, e);
    const double g = e * pow(a.v, e - 1.0);
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = g * a.d[k];
    return r;
}

template <int K> __device__ __forceinline__ Dual<K> fast_rcp(const Dual<K>& a) {
    Dual<K> r;
    r.v = 1.0 / a.v;
    const double g = -r.v * r.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = g * a.d[k];
    return r;
}

// u at a cv-edge midpoint, u = alpha x + beta y + gamma (src/problem.jl:428-434): the three terms are ~|u| |x| / h and
// cancel to |u|, so the reference's individually rounded (alpha x + beta y) + gamma is kept -- a contracted FMA differs
// by an ulp of the big terms, 1e-12 of u on the 4096^2 lattice
__device__ __forceinline__ double shape_value(double a, double b, double g, double x, double y) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), g);
}
template <int K>
__device__ __forceinline__ Dual<K> shape_value(const Dual<K>& a, const Dual<K>& b, const Dual<K>& g, double x, double y) {
    return a * x + b * y + g;
}
// s_a u_i + s_b u_j + s_c u_k (shape_functions.jl:2-19) with the reference's roundings, for the same reason
__device__ __forceinline__ double shape_coeff(double sa, double sb, double sc, double ui, double uj, double uk) {
    return __dadd_rn(__dadd_rn(__dmul_rn(sa, ui), __dmul_rn(sb, uj)), __dmul_rn(sc, uk));
}

// ---- flux registry: q(x, y, t, alpha, beta, gamma, p), src/problem.jl:113-116, 425-440 ----
template <int MODEL, int NEQ, class T>
__device__ __forceinline__ void flux_eval(const FluxParams& fp, double x, double y, double t, const T* a, const T* b, const T* g,
                                          double dtab, T* qx, T* qy) {
    if constexpr (MODEL == FVM_FLUX_DIFF_CONST) {
#pragma unroll
        for (int v = 0; v < NEQ; ++v) {
            qx[v] = -fp.p[v] * a[v];
            qy[v] = -fp.p[v] * b[v];
        }
    } else if constexpr (MODEL == FVM_FLUX_DIFF_TABLE) {
#pragma unroll
        for (int v = 0; v < NEQ; ++v) {
            qx[v] = -dtab * a[v];
            qy[v] = -dtab * b[v];
        }
    } else if constexpr (MODEL == FVM_FLUX_DIFF_POWER) {
#pragma unroll
        for (int v = 0; v < NEQ; ++v) {
            const T u = shape_value(a[v], b[v], g[v], x, y);
            const double D0 = fp.p[3 * v], mm = fp.p[3 * v + 1];
            const T base = fp.p[3 * v + 2] != 0.0 ? fabs_t(u) : u;
            T D;
            if (mm == 1.0) D = T(D0);
            else if (mm == 2.0) D = D0 * base;
            else if (mm == 3.0) D = D0 * base * base;
            else D = D0 * pow_t(base, mm - 1.0);
            qx[v] = -(D * a[v]);
            qy[v] = -(D * b[v]);
        }
    } else if constexpr (MODEL == FVM_FLUX_ADVDIFF) {
#pragma unroll
        for (int v = 0; v < NEQ; ++v) {
            const T u = shape_value(a[v], b[v], g[v], x, y);
            const double D = fp.p[3 * v];
            qx[v] = fp.p[3 * v + 1] * u - D * a[v];
            qy[v] = fp.p[3 * v + 2] * u - D * b[v];
        }
    } else if constexpr (MODEL == FVM_FLUX_KELLER_SEGEL) {
        // src/FiniteVolumeMethod.jl:98-110 : q_u = chi(u) grad v - grad u ; q_v = -D grad v
        static_assert(NEQ == 2 || MODEL != FVM_FLUX_KELLER_SEGEL, "Keller-Segel is a 2-species model");
        const T u = shape_value(a[0], b[0], g[0], x, y);
        const T chi = fp.p[0] * u * fast_rcp(u * u + 1.0);
        qx[0] = chi * a[1] - a[0];
        qy[0] = chi * b[1] - b[0];
        qx[1] = -fp.p[1] * a[1];
        qy[1] = -fp.p[1] * b[1];
    }
}

template <int MODEL>
struct FluxTraits {
    // does the flux read (x, y, gamma)?  If not, only s1..s6 and the scaled normals are streamed.
    static constexpr bool full = (MODEL == FVM_FLUX_DIFF_POWER || MODEL == FVM_FLUX_ADVDIFF || MODEL == FVM_FLUX_KELLER_SEGEL);
    static constexpr bool table = (MODEL == FVM_FLUX_DIFF_TABLE);
};

// ---- sources S(x, y, t, u, p), src/problem.jl:10-13, 342-345 -------------------------------
template <int NEQ, class T>
__device__ __forceinline__ T source_eval_t(const SourceParams& sp, int v, const T* u, const double* tab) {
    switch (sp.model) {
        case FVM_SRC_LINEAR: return sp.p[2 * v] * u[v] + sp.p[2 * v + 1];
        case FVM_SRC_LOGISTIC: return sp.p[v] * u[v] * (1.0 - u[v]);
        case FVM_SRC_TABLE: return T(tab[v]);
        case FVM_SRC_GRAY_SCOTT:
            if constexpr (NEQ >= 2) return v == 0 ? sp.p[0] * (1.0 - u[0]) - u[0] * (u[1] * u[1]) : -(sp.p[1] * u[1]) + u[0] * (u[1] * u[1]);
            return T(0.0);
        case FVM_SRC_BRUSSELATOR:
            if constexpr (NEQ >= 2) return v == 0 ? (u[0] * u[0]) * u[1] - 2.0 * u[0] : -((u[0] * u[0]) * u[1]) + u[0];
            return T(0.0);
        case FVM_SRC_KELLER_SEGEL:
            if constexpr (NEQ >= 2) return v == 0 ? u[0] * (1.0 - u[0]) : u[0] - sp.p[0] * u[1];
            return T(0.0);
        default: return T(0.0);  // zero(eltype(u)), src/problem.jl:123
    }
}
template <int NEQ>
__device__ __forceinline__ double source_eval(const SourceParams& sp, int v, const double* u, const double* tab) {
    return source_eval_t<NEQ, double>(sp, v, u, tab);
}

// ---- condition functions a(x, y, t, u, p), src/conditions.jl:16-20 -------------------------
template <class T>
__device__ __forceinline__ T cond_eval_t(const CondFn& c, double x, double y, double t, const T& u) {
    switch (c.id) {
        case FVM_COND_AFFINE_U: return c.p[1] * u + c.p[0];
        case FVM_COND_EXP_SAT: return T(c.p[0] * (1.0 - exp(-t / c.p[1])));
        case FVM_COND_LINEAR_XY: return T(c.p[0] + c.p[1] * x + c.p[2] * y);
        case FVM_COND_EXP_XYT: return T(c.p[0] * exp(c.p[1] * x + c.p[2] * y + c.p[3] * t));
        default: return T(c.p[0]);
    }
}
__device__ __forceinline__ double cond_eval(const CondFn& c, double x, double y, double t, double u) {
    switch (c.id) {
        case FVM_COND_AFFINE_U: return c.p[0] + c.p[1] * u;
        case FVM_COND_EXP_SAT: return c.p[0] * (1.0 - exp(-t / c.p[1]));
        case FVM_COND_LINEAR_XY: return c.p[0] + c.p[1] * x + c.p[2] * y;
        case FVM_COND_EXP_XYT: return c.p[0] * exp(c.p[1] * x + c.p[2] * y + c.p[3] * t);
        default: return c.p[0];
    }
}

// ---- node pass, src/equations/source_contributions.jl:33-68 --------------------------------
template <int NEQ>
__device__ __forceinline__ void node_finish(const DevMesh& m, const SourceParams& sp, double t, int g, const double* acc,
                                            const double* uv, double* __restrict__ du, const double V, const uint8_t* kinds) {
    double tab[NEQ];
    if (sp.model == FVM_SRC_TABLE) {
#pragma unroll
        for (int v = 0; v < NEQ; ++v) tab[v] = m.src_tab[(size_t)g * NEQ + v];
    }
#pragma unroll
    for (int v = 0; v < NEQ; ++v) {
        const uint8_t kind = kinds[v];
        double out;
        if (kind == FVM_NODE_FREE) {
            out = acc[v] / V + source_eval<NEQ>(sp, v, uv, tab);
        } else if (kind == FVM_NODE_DUDT) {
            const CondFn c = m.cond[v * FVM_MAX_COND_FN + m.fidx[(size_t)v * m.n_nodes + g]];
            out = cond_eval(c, m.xy[2 * (size_t)g], m.xy[2 * (size_t)g + 1], t, uv[v]);
        } else {
            out = 0.0;  // Dirichlet node, or a point that is not a vertex
        }
        du[(size_t)g * NEQ + v] = out;
    }
}

template <int NEQ>
__device__ __forceinline__ void node_finish(const DevMesh& m, const SourceParams& sp, double t, int g, const double* acc,
                                            const double* uv, double* __restrict__ du) {
    uint8_t kinds[NEQ];
#pragma unroll
    for (int v = 0; v < NEQ; ++v) kinds[v] = m.kind[(size_t)v * m.n_nodes + g];
    node_finish<NEQ>(m, sp, t, g, acc, uv, du, m.vol[g], kinds);
}
