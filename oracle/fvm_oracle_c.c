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
