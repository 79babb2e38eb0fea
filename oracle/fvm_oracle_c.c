/*
 * oracle/fvm_oracle_c.c -- CPU restatement of the reference's RHS in C.  TEST INFRASTRUCTURE ONLY.
 *
 * Used (a) by tests/ to cross-check the NumPy oracle and (b) by bench.py's `cpu_baseline` and
 * `--impl reference` legs as the timed CPU arm ("port": Julia is not installed).
 * The product never links or calls this file.
 *
 * It keeps the STRUCTURE of the reference (so the baseline is not flattered or handicapped):
 *   - triangle properties live in a hash table keyed by the vertex triple, looked up once per
 *     triangle per call (Dict{NTuple{3,Int},TriangleProperties}, src/geometry.jl:44-49,61-65);
 *   - node conditions are looked up in hash sets (src/conditions.jl:310-316,342-484);
 *   - the threaded path splits triangles into nthreads contiguous chunks, each thread scatters
 *     into its own full copy of du, and the copies are combined serially
 *     (src/solve.jl:1-27, src/equations/main_equations.jl:47-81, triangle_contributions.jl:46-70);
 *   - serial order: zero du, triangles, boundary edges, nodes (main_equations.jl:38-44).
 * Flux: the diffusion form q = -D(x,y,t,u)(alpha,beta) of src/problem.jl:425-440 with constant D
 * (README config), source 0, Dirichlet value 0 -- the workload bench.py measures.
 * A second entry point runs the same arithmetic on flat arrays without the hash containers
 * ("fair CPU line").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double s[9];
    double mid[6];
    double nrm[6];
    double len[3];
} TriProps; /* TriangleProperties, src/geometry.jl:21-26 */

typedef struct {
    int64_t N, T, Eb;
    const double* xy;
    const int32_t* tri;
    double* vol;
    /* Dict{NTuple{3,Int},TriangleProperties} as an open-addressing table */
    uint64_t cap;
    int64_t* keys; /* 3 per slot, -1 = empty */
    TriProps* props;
    /* Dict{Int,Int} dirichlet_nodes as an open-addressing set */
    uint64_t dcap;
    int64_t* dkeys;
    /* flat copy in triangle order for the "fair" variant */
    TriProps* flat;
    uint8_t* is_dir;
    double D;
    int nthreads;
    double* dup; /* N x nthreads */
} Oracle;

static inline uint64_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static inline uint64_t hash3(int64_t i, int64_t j, int64_t k) { return mix(mix(mix((uint64_t)i) ^ (uint64_t)j) ^ (uint64_t)k); }

static void tri_props(const double* xy, int32_t i, int32_t j, int32_t k, TriProps* P, double* S) {
    /* src/geometry.jl:107-161 */
    double px = xy[2 * i], py = xy[2 * i + 1], qx = xy[2 * j], qy = xy[2 * j + 1], rx = xy[2 * k], ry = xy[2 * k + 1];
    double cx = (px + qx + rx) / 3, cy = (py + qy + ry) / 3;
    double m1x = (px + qx) / 2, m1y = (py + qy) / 2, m2x = (qx + rx) / 2, m2y = (qy + ry) / 2, m3x = (rx + px) / 2,
           m3y = (ry + py) / 2;
    S[0] = 0.5 * fabs((cx - px) * (m1y - m3y) - (cy - py) * (m1x - m3x));
    S[1] = 0.5 * fabs((cx - qx) * (m2y - m1y) - (cy - qy) * (m2x - m1x));
    S[2] = 0.5 * fabs((cx - rx) * (m3y - m2y) - (cy - ry) * (m3x - m2x));
    double Dl = qx * ry - qy * rx - px * ry + rx * py + px * qy - qx * py;
    P->s[0] = (qy - ry) / Dl;
    P->s[1] = (ry - py) / Dl;
    P->s[2] = (py - qy) / Dl;
    P->s[3] = (rx - qx) / Dl;
    P->s[4] = (px - rx) / Dl;
    P->s[5] = (qx - px) / Dl;
    P->s[6] = (qx * ry - rx * qy) / Dl;
    P->s[7] = (rx * py - px * ry) / Dl;
    P->s[8] = (px * qy - qx * py) / Dl;
    double mx[3] = {m1x, m2x, m3x}, my[3] = {m1y, m2y, m3y};
    for (int e = 0; e < 3; ++e) {
        P->mid[2 * e] = (mx[e] + cx) / 2;
        P->mid[2 * e + 1] = (my[e] + cy) / 2;
        double ex = cx - mx[e], ey = cy - my[e];
        double l = sqrt(ex * ex + ey * ey);
        P->len[e] = l;
        P->nrm[2 * e] = ey / l;
        P->nrm[2 * e + 1] = -ex / l;
    }
}

void* oracle_create(const double* xy, int64_t N, const int32_t* tri, int64_t T, const int32_t* dir_nodes, int64_t n_dir,
                    double D, int nthreads) {
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->N = N;
    o->T = T;
    o->xy = xy;
    o->tri = tri;
    o->D = D;
    o->nthreads = nthreads > 0 ? nthreads : 1;
    o->vol = (double*)calloc(N, sizeof(double));
    o->cap = 1;
    while (o->cap < (uint64_t)(2 * T)) o->cap <<= 1;
    o->keys = (int64_t*)malloc(sizeof(int64_t) * 3 * o->cap);
    for (uint64_t s = 0; s < 3 * o->cap; ++s) o->keys[s] = -1;
    o->props = (TriProps*)malloc(sizeof(TriProps) * o->cap);
    o->flat = (TriProps*)malloc(sizeof(TriProps) * T);
    for (int64_t t = 0; t < T; ++t) { /* serial, like FVMGeometry(tri) */
        int32_t i = tri[3 * t], j = tri[3 * t + 1], k = tri[3 * t + 2];
        TriProps P;
        double S[3];
        tri_props(xy, i, j, k, &P, S);
        o->vol[i] += S[0];
        o->vol[j] += S[1];
        o->vol[k] += S[2];
        uint64_t s = hash3(i, j, k) & (o->cap - 1);
        while (o->keys[3 * s] >= 0) s = (s + 1) & (o->cap - 1);
        o->keys[3 * s] = i;
        o->keys[3 * s + 1] = j;
        o->keys[3 * s + 2] = k;
        o->props[s] = P;
        o->flat[t] = P;
    }
    o->dcap = 16;
    while (o->dcap < (uint64_t)(2 * n_dir + 1)) o->dcap <<= 1;
    o->dkeys = (int64_t*)malloc(sizeof(int64_t) * o->dcap);
    for (uint64_t s = 0; s < o->dcap; ++s) o->dkeys[s] = -1;
    o->is_dir = (uint8_t*)calloc(N, 1);
    for (int64_t q = 0; q < n_dir; ++q) {
        uint64_t s = mix((uint64_t)dir_nodes[q]) & (o->dcap - 1);
        while (o->dkeys[s] >= 0 && o->dkeys[s] != dir_nodes[q]) s = (s + 1) & (o->dcap - 1);
        o->dkeys[s] = dir_nodes[q];
        o->is_dir[dir_nodes[q]] = 1;
    }
    o->dup = (double*)malloc(sizeof(double) * N * o->nthreads);
    return o;
}

void oracle_destroy(void* p) {
    Oracle* o = (Oracle*)p;
    free(o->vol);
    free(o->keys);
    free(o->props);
    free(o->flat);
    free(o->dkeys);
    free(o->is_dir);
    free(o->dup);
    free(o);
}

void oracle_volumes(void* p, double* out) { memcpy(out, ((Oracle*)p)->vol, sizeof(double) * ((Oracle*)p)->N); }

static inline const TriProps* get_triangle_props(const Oracle* o, int64_t i, int64_t j, int64_t k) {
    uint64_t s = hash3(i, j, k) & (o->cap - 1);
    while (!(o->keys[3 * s] == i && o->keys[3 * s + 1] == j && o->keys[3 * s + 2] == k)) s = (s + 1) & (o->cap - 1);
    return &o->props[s];
}
static inline int is_dirichlet_node(const Oracle* o, int64_t i) {
    uint64_t s = mix((uint64_t)i) & (o->dcap - 1);
    while (o->dkeys[s] >= 0) {
        if (o->dkeys[s] == i) return 1;
        s = (s + 1) & (o->dcap - 1);
    }
    return 0;
}

/* fvm_eqs_single_triangle!, src/equations/triangle_contributions.jl:28-35 */
static inline void single_triangle(double* du, const double* u, const TriProps* P, int32_t i, int32_t j, int32_t k, double D) {
    double a = P->s[0] * u[i] + P->s[1] * u[j] + P->s[2] * u[k];
    double b = P->s[3] * u[i] + P->s[4] * u[j] + P->s[5] * u[k];
    double g = P->s[6] * u[i] + P->s[7] * u[j] + P->s[8] * u[k];
    double Q[3];
    for (int e = 0; e < 3; ++e) {
        double x = P->mid[2 * e], y = P->mid[2 * e + 1];
        double uu = a * x + b * y + g; /* construct_flux_function, src/problem.jl:428-434 */
        (void)uu;
        double qx = -D * a, qy = -D * b;
        Q[e] = (qx * P->nrm[2 * e] + qy * P->nrm[2 * e + 1]) * P->len[e];
    }
    du[i] = du[i] + Q[2] - Q[0];
    du[j] = du[j] + Q[0] - Q[1];
    du[k] = du[k] + Q[1] - Q[2];
}

static void node_pass(const Oracle* o, double* du, int use_hash) {
    /* src/equations/source_contributions.jl:33-47; all-Dirichlet or free nodes, S = 0 */
#pragma omp parallel for schedule(static) num_threads(o->nthreads)
    for (int64_t i = 0; i < o->N; ++i) {
        int dir = use_hash ? is_dirichlet_node(o, i) : o->is_dir[i];
        du[i] = dir ? 0.0 : du[i] / o->vol[i] + 0.0;
    }
}

/* serial_fvm_eqs!, main_equations.jl:38-44 (boundary edges: dead work on all-Dirichlet, D-6) */
void oracle_fvm_eqs_serial(void* p, const double* u, double* du) {
    Oracle* o = (Oracle*)p;
    memset(du, 0, sizeof(double) * o->N);
    for (int64_t t = 0; t < o->T; ++t) {
        int32_t i = o->tri[3 * t], j = o->tri[3 * t + 1], k = o->tri[3 * t + 2];
        single_triangle(du, u, get_triangle_props(o, i, j, k), i, j, k, o->D);
    }
    int nt = o->nthreads;
    o->nthreads = 1;
    node_pass(o, du, 1);
    o->nthreads = nt;
}

/* parallel_fvm_eqs!, main_equations.jl:47-81 */
void oracle_fvm_eqs_threaded(void* p, const double* u, double* du) {
    Oracle* o = (Oracle*)p;
    const int nt = o->nthreads;
    memset(du, 0, sizeof(double) * o->N);
    memset(o->dup, 0, sizeof(double) * o->N * nt); /* fill!(_duplicated_du, 0) */
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int c = omp_get_thread_num();
#else
        int c = 0;
#endif
        int64_t lo = o->T * c / nt, hi = o->T * (c + 1) / nt; /* index_chunks(solid_triangles; n = nt) */
        double* mine = o->dup + (size_t)o->N * c;
        for (int64_t t = lo; t < hi; ++t) {
            int32_t i = o->tri[3 * t], j = o->tri[3 * t + 1], k = o->tri[3 * t + 2];
            single_triangle(mine, u, get_triangle_props(o, i, j, k), i, j, k, o->D);
        }
    }
    for (int c = 0; c < nt; ++c) { /* combine_duplicated_du!, serial */
        const double* col = o->dup + (size_t)o->N * c;
        for (int64_t i = 0; i < o->N; ++i) du[i] += col[i];
    }
    node_pass(o, du, 1);
}

/* fair CPU line: flat SoA-ish arrays, no hash containers, same arithmetic, owner chunks */
void oracle_fvm_eqs_flat(void* p, const double* u, double* du) {
    Oracle* o = (Oracle*)p;
    const int nt = o->nthreads;
    memset(du, 0, sizeof(double) * o->N);
    memset(o->dup, 0, sizeof(double) * o->N * nt);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int c = omp_get_thread_num();
#else
        int c = 0;
#endif
        int64_t lo = o->T * c / nt, hi = o->T * (c + 1) / nt;
        double* mine = o->dup + (size_t)o->N * c;
        for (int64_t t = lo; t < hi; ++t) {
            int32_t i = o->tri[3 * t], j = o->tri[3 * t + 1], k = o->tri[3 * t + 2];
            single_triangle(mine, u, &o->flat[t], i, j, k, o->D);
        }
    }
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < o->N; ++i) {
        double acc = 0.0;
        for (int c = 0; c < nt; ++c) acc += o->dup[(size_t)o->N * c + i];
        du[i] = acc;
    }
    node_pass(o, du, 0);
}

/* ------------------------------------------------------------------------------------------------
 * General stateless RHS for the full-size parity tests (BASELINE configs 2 and 4 at 4096^2): scalar problems and
 * FVMSystems with the registered flux / source forms, geometry recomputed per triangle (no Dict: 33.5 M
 * TriangleProperties would need 13 GB), serial, reference order:
 *   zero du; triangles in the given order (fvm_eqs_single_triangle!, triangle_contributions.jl:10-35, with
 *   get_shape_function_coefficients shape_functions.jl:2-19 and get_flux individual_flux_contributions.jl:27-50);
 *   node pass (source_contributions.jl:33-68).
 * The boundary-edge pass (boundary_edge_contributions.jl:77-86) is omitted: this entry point is only valid where it
 * adds nothing -- boundary nodes all Dirichlet (their rows are overwritten in the node pass) or homogeneous Neumann
 * edges (a = 0, so du[i] -= 0 exactly).  The caller states which nodes are Dirichlet per species.
 * Flux forms (the tutorial closures the device registry mirrors; u = alpha x + beta y + gamma at the cv-edge midpoint):
 *   0 constant D_v:            q_v = (-D_v alpha_v, -D_v beta_v)                          problem.jl:425-440
 *   2 power law:               D = D0_v |u_v|^(m_v - 1) (|.| iff flag), q_v = (-D alpha_v, -D beta_v)
 *   3 advection-diffusion:     q_v = (nux_v u_v - D_v alpha_v, nuy_v u_v - D_v beta_v)
 *   4 Keller-Segel (2 species): chi = c u/(1 + u^2); q_u = (chi alpha_2 - alpha_1, chi beta_2 - beta_1); q_v = -D grad v
 *                                                                          src/FiniteVolumeMethod.jl:98-110
 * Sources: 0 zero; 1 lam_v u_v + mu_v; 2 lam_v u_v (1 - u_v); 4 Gray-Scott; 5 Brusselator; 6 Keller-Segel (u(1-u); u - a v).
 * u and du are species-interleaved per node (Julia's column-major Matrix(neq, N)).
 */
static void general_flux(int model, const double* fp, int neq, double x, double y, const double* a, const double* b, const double* g,
                         double* qx, double* qy) {
    for (int v = 0; v < neq; ++v) {
        if (model == 0) {
            qx[v] = -fp[v] * a[v];
            qy[v] = -fp[v] * b[v];
        } else if (model == 2) {
            double u = a[v] * x + b[v] * y + g[v];
            double base = fp[3 * v + 2] != 0.0 ? fabs(u) : u;
            double D = fp[3 * v] * pow(base, fp[3 * v + 1] - 1.0);
            qx[v] = -D * a[v];
            qy[v] = -D * b[v];
        } else if (model == 3) {
            double u = a[v] * x + b[v] * y + g[v];
            qx[v] = fp[3 * v + 1] * u - fp[3 * v] * a[v];
            qy[v] = fp[3 * v + 2] * u - fp[3 * v] * b[v];
        }
    }
    if (model == 4) {
        double u = a[0] * x + b[0] * y + g[0];
        double chi = fp[0] * u / (1 + u * u);
        qx[0] = chi * a[1] - a[0];
        qy[0] = chi * b[1] - b[0];
        qx[1] = -fp[1] * a[1];
        qy[1] = -fp[1] * b[1];
    }
}

static double general_source(int model, const double* sp, int v, const double* u) {
    switch (model) {
        case 1: return sp[2 * v] * u[v] + sp[2 * v + 1];
        case 2: return sp[v] * u[v] * (1 - u[v]);
        case 4: return v == 0 ? sp[0] * (1 - u[0]) - u[0] * (u[1] * u[1]) : -sp[1] * u[1] + u[0] * (u[1] * u[1]);
        case 5: return v == 0 ? (u[0] * u[0]) * u[1] - 2 * u[0] : -((u[0] * u[0]) * u[1]) + u[0];
        case 6: return v == 0 ? u[0] * (1 - u[0]) : u[0] - sp[0] * u[1];
        default: return 0.0;
    }
}

int oracle_fvm_eqs_general(const double* xy, int64_t N, const int32_t* tri, int64_t T, int neq, int flux_model, const double* fp,
                           int src_model, const double* sp, const uint8_t* is_dirichlet /* [neq][N] */, const double* u, double* du) {
    if (neq < 1 || neq > 4 || (flux_model == 4 && neq != 2)) return 1;
    double* vol = (double*)calloc(N, sizeof(double));
    uint8_t* is_vertex = (uint8_t*)calloc(N, 1);
    memset(du, 0, sizeof(double) * N * neq);
    for (int64_t t = 0; t < T; ++t) {
        const int32_t i = tri[3 * t], j = tri[3 * t + 1], k = tri[3 * t + 2];
        TriProps P;
        double S[3];
        tri_props(xy, i, j, k, &P, S);
        vol[i] += S[0]; /* geometry.jl:126-135 */
        vol[j] += S[1];
        vol[k] += S[2];
        is_vertex[i] = is_vertex[j] = is_vertex[k] = 1;
        double a[4], b[4], g[4], Q[3][4];
        for (int v = 0; v < neq; ++v) { /* shape_functions.jl:2-19 */
            const double ui = u[(size_t)i * neq + v], uj = u[(size_t)j * neq + v], uk = u[(size_t)k * neq + v];
            a[v] = P.s[0] * ui + P.s[1] * uj + P.s[2] * uk;
            b[v] = P.s[3] * ui + P.s[4] * uj + P.s[5] * uk;
            g[v] = P.s[6] * ui + P.s[7] * uj + P.s[8] * uk;
        }
        for (int e = 0; e < 3; ++e) { /* individual_flux_contributions.jl:21-50 */
            double qx[4], qy[4];
            general_flux(flux_model, fp, neq, P.mid[2 * e], P.mid[2 * e + 1], a, b, g, qx, qy);
            for (int v = 0; v < neq; ++v) Q[e][v] = (qx[v] * P.nrm[2 * e] + qy[v] * P.nrm[2 * e + 1]) * P.len[e];
        }
        for (int v = 0; v < neq; ++v) { /* update_du!, triangle_contributions.jl:10-25 */
            du[(size_t)i * neq + v] = du[(size_t)i * neq + v] + Q[2][v] - Q[0][v];
            du[(size_t)j * neq + v] = du[(size_t)j * neq + v] + Q[0][v] - Q[1][v];
            du[(size_t)k * neq + v] = du[(size_t)k * neq + v] + Q[1][v] - Q[2][v];
        }
    }
    for (int64_t i = 0; i < N; ++i) /* source_contributions.jl:33-68 */
        for (int v = 0; v < neq; ++v) {
            double* d = du + (size_t)i * neq + v;
            if (!is_vertex[i] || (is_dirichlet && is_dirichlet[(size_t)v * N + i])) *d = 0.0;
            else *d = *d / vol[i] + general_source(src_model, sp, v, u + (size_t)i * neq);
        }
    free(vol);
    free(is_vertex);
    return 0;
}

/* CSC-free CSR y = A x + b, single-threaded like SparseArrays' mul! (diffusion_equation.jl:93-94) */
void oracle_spmv(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val, const double* b, const double* x,
                 double* y, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int32_t q = rowptr[i]; q < rowptr[i + 1]; ++q) acc += val[q] * x[col[q]];
        y[i] = acc + (b ? b[i] : 0.0);
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
