/* The README problem of FiniteVolumeMethod.jl (README.md:20-45: 50x50 lattice on [0,2]^2, D = 1/9, Dirichlet u = 0)
 * through the C ABI of libfvmcuda.so, from plain C99 — what a Julia `ccall` binding does, without Julia:
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include examples/readme_diffusion.c \
 *       -L finitevolumemethod.jl_b200/lib -lfvmcuda -Wl,-rpath,$PWD/finitevolumemethod.jl_b200/lib -lm -o readme_diffusion
 *
 * The host-only entry points (plan self-check, graph partition, FVMWIRE container) run anywhere; the compute
 * part needs a B200 and exits with status 2 and the library's "no CPU fallback" message without one. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fvmcuda.h"

#define NX 50
#define NY 50
#define N (NX * NY)
#define T (2 * (NX - 1) * (NY - 1))
#define EB (2 * (NX - 1) + 2 * (NY - 1))

static int fail(const char* what, fvm_handle h) {
    fprintf(stderr, "%s: %s\n", what, fvm_last_error(h));
    return 1;
}

int main(void) {
    static double xy[2 * N], u[N], du[N];
    static int32_t tri[3 * T], edges[2 * EB], owner[N];
    static uint8_t kind[N];
    static int32_t fidx[N];
    int64_t stats[8], cut = 0;
    int i, j, k = 0, e = 0;
    /* triangulate_rectangle(0, 2, 0, 2, 50, 50; single_boundary = true), 1-based like Julia */
    for (j = 0; j < NY; ++j)
        for (i = 0; i < NX; ++i) {
            xy[2 * (i + j * NX)] = 2.0 * i / (NX - 1);
            xy[2 * (i + j * NX) + 1] = 2.0 * j / (NY - 1);
        }
    for (j = 0; j < NY - 1; ++j)
        for (i = 0; i < NX - 1; ++i) {
            const int32_t p = 1 + i + j * NX;
            tri[k++] = p; tri[k++] = p + 1; tri[k++] = p + NX;
            tri[k++] = p + NX; tri[k++] = p + 1; tri[k++] = p + NX + 1;
        }
    for (i = 0; i < NX - 1; ++i) { edges[e++] = 1 + i; edges[e++] = 2 + i; }                                        /* bottom */
    for (j = 0; j < NY - 1; ++j) { edges[e++] = NX + j * NX; edges[e++] = NX + (j + 1) * NX; }                      /* right  */
    for (i = NX - 1; i > 0; --i) { edges[e++] = 1 + i + (NY - 1) * NX; edges[e++] = i + (NY - 1) * NX; }            /* top    */
    for (j = NY - 1; j > 0; --j) { edges[e++] = 1 + j * NX; edges[e++] = 1 + (j - 1) * NX; }                        /* left   */
    memset(kind, 0, sizeof kind);
    memset(fidx, 0, sizeof fidx);
    for (i = 0; i < 2 * EB; ++i) kind[edges[i] - 1] = FVM_NODE_DIRICHLET;

    /* ---- host-only: plan self-check, partition, wire container -------------------------------------- */
    if (fvm_plan_selftest(xy, N, tri, T, 1, 1, edges, EB, kind, 256, stats)) return fail("fvm_plan_selftest", NULL);
    printf("plan: %lld tiles, %lld vertices, %lld interface nodes, %lld live boundary edges\n", (long long)stats[0],
           (long long)stats[1], (long long)stats[2], (long long)stats[6]);
    if (fvm_partition_graph(N, tri, T, 1, 4, owner) || fvm_partition_edge_cut(N, tri, T, 1, owner, &cut))
        return fail("fvm_partition_graph", NULL);
    {
        int cnt[4] = {0, 0, 0, 0};
        for (i = 0; i < N; ++i) cnt[owner[i]]++;
        printf("partition: %d %d %d %d nodes, edge cut %lld\n", cnt[0], cnt[1], cnt[2], cnt[3], (long long)cut);
    }
    {
        const char* path = "readme_mesh.fvmw";
        fvm_wire_handle w = NULL;
        const int64_t dp[2] = {2, N}, dt[2] = {3, T}, de[2] = {2, EB}, d1[1] = {1};
        const int32_t base = 1;
        int32_t idx = -1;
        static double back[2 * N];
        if (fvm_wire_create(path, &w) || fvm_wire_put(w, "points", FVM_WIRE_F64, 2, dp, xy) ||
            fvm_wire_put(w, "triangles", FVM_WIRE_I32, 2, dt, tri) || fvm_wire_put(w, "index_base", FVM_WIRE_I32, 1, d1, &base) ||
            fvm_wire_put(w, "boundary_edges", FVM_WIRE_I32, 2, de, edges) || fvm_wire_close(w)) {
            fprintf(stderr, "wire write: %s\n", fvm_wire_last_error(NULL));
            return 1;
        }
        if (fvm_wire_open(path, &w) || fvm_wire_find(w, "points", &idx) || fvm_wire_get(w, idx, back, (int64_t)sizeof back) ||
            fvm_wire_close(w) || memcmp(back, xy, sizeof back)) {
            fprintf(stderr, "wire read: %s\n", fvm_wire_last_error(NULL));
            return 1;
        }
        printf("wire: %s written and verified (crc32 of points %08x)\n", path, (unsigned)fvm_wire_crc32(xy, (int64_t)sizeof xy));
        {   /* ---- compute: needs a CUDA device ------------------------------------------------------------ */
            fvm_handle h = NULL;
            const double zero = 0.0, D = 1.0 / 9.0;
            double amax = 0.0;
            int32_t rc = fvm_create_from_wire(path, 1, 0, &h);
            remove(path);
            if (rc == FVM_ERR_CUDA) {
                fprintf(stderr, "%s\n", fvm_last_error(NULL));
                return 2;
            }
            if (rc) return fail("fvm_create_from_wire", NULL);
            if (fvm_set_node_conditions(h, 0, kind, fidx) || fvm_set_condition_fn(h, 0, 0, FVM_COND_CONST, &zero, 1) ||
                fvm_set_flux(h, FVM_FLUX_DIFF_CONST, &D, 1) || fvm_finalize(h, 0, 0))
                return fail("setup", h);
            for (i = 0; i < N; ++i) u[i] = xy[2 * i + 1] <= 1.0 ? 50.0 : 0.0; /* f(x, y) of the README */
            if (fvm_rhs(h, 0.0, u, du, 0)) return fail("fvm_rhs", h);
            for (i = 0; i < N; ++i) amax = fmax(amax, fabs(du[i]));
            printf("fvm_eqs!: max |du| = %.6e at t = 0\n", amax);
            if (fvm_tsit5(h, 0, u, 0.0, 0.5, 0.0025, 0, NULL, NULL, 0)) return fail("fvm_tsit5", h);
            for (amax = 0.0, i = 0; i < N; ++i) amax = fmax(amax, u[i]);
            printf("Tsit5 to t = 0.5: max u = %.6f\n", amax);
            fvm_destroy(h);
        }
    }
    return 0;
}
