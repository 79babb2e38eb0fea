/*
 * fvmcuda.h — C ABI of libfvmcuda.so, the B200 (sm_100a) engine behind
 * FiniteVolumeMethod.jl's semi-discrete right-hand side `fvm_eqs!` and its
 * linear-template operator path.
 *
 * Conventions
 *  - every call returns an int32 status (FVM_OK == 0); fvm_last_error(h) gives
 *    the message of the last failing call on that handle;
 *  - the caller owns every host buffer, the library owns every device buffer
 *    behind the handle; a handle is bound to one device and one stream and is
 *    not thread-safe; distinct handles may be used from distinct host threads;
 *  - vectors cross the ABI in the CALLER's node/triangle numbering and in the
 *    reference's memory layout: a system state `u::Matrix(neq, N)` (column
 *    major) is species-interleaved per node.  The tile/Hilbert renumbering is
 *    internal ("native" order); the *_native entry points expose it so an
 *    integrator can keep its state resident without a permutation per call;
 *  - `on_device` != 0 means the pointer arguments are device pointers on the
 *    handle's device;
 *  - there is no CPU fallback: without a CUDA device every compute call fails
 *    with FVM_ERR_CUDA.
 *
 * File:line citations are relative to /root/reference (FiniteVolumeMethod.jl v1.2.3).
 */
#ifndef FVMCUDA_H
#define FVMCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fvm_ctx* fvm_handle;

enum fvm_status {
    FVM_OK = 0,
    FVM_ERR_ARG = 1,         /* invalid argument (the reference's @assert / ArgumentError) */
    FVM_ERR_CUDA = 2,        /* CUDA runtime failure, including "no device" */
    FVM_ERR_UNSUPPORTED = 3, /* functor / closure outside the registry (north_star b) */
    FVM_ERR_STATE = 4,       /* call order violated (e.g. rhs before finalize) */
    FVM_ERR_NCCL = 5,
    FVM_ERR_IO = 6           /* FVMWIRE container: file missing, truncated, corrupt or checksum mismatch */
};

/* node condition kinds: src/conditions.jl:36-41, flattened per node / per boundary edge */
enum fvm_node_kind { FVM_NODE_FREE = 0, FVM_NODE_DIRICHLET = 1, FVM_NODE_DUDT = 2 };
enum fvm_edge_kind { FVM_EDGE_NONE = 0, FVM_EDGE_NEUMANN = 1, FVM_EDGE_CONSTRAINED = 2 };

/* Flux models q(x,y,t,alpha,beta,gamma,p): the registry that replaces Julia closures
 * (src/problem.jl:113-116, 425-440).  One model per handle; parameters are per species. */
enum fvm_flux_model {
    FVM_FLUX_DIFF_CONST = 0,  /* q_v = -D_v grad u_v                 params: D_v (neq)            */
    FVM_FLUX_DIFF_TABLE = 1,  /* q_v = -D(x_e) grad u_v, D tabulated by fvm_set_flux_table        */
    FVM_FLUX_DIFF_POWER = 2,  /* D = D0_v |u_v|^(m_v-1)              params: (D0_v, m_v) per v    */
    FVM_FLUX_ADVDIFF = 3,     /* q_v = nu_v u_v - D_v grad u_v       params: (D_v, nux_v, nuy_v)  */
    FVM_FLUX_KELLER_SEGEL = 4 /* neq=2: q_u = chi(u) grad v - grad u, chi = c u/(1+u^2);
                                 q_v = -D grad v                    params: (c, D)               */
};

/* Source models S(x,y,t,u,p): src/problem.jl:10-13, 342-345 */
enum fvm_source_model {
    FVM_SRC_ZERO = 0,
    FVM_SRC_LINEAR = 1,      /* S_v = lam_v u_v + mu_v              params: (lam_v, mu_v) per v   */
    FVM_SRC_LOGISTIC = 2,    /* S_v = lam_v u_v (1 - u_v)           params: lam_v                 */
    FVM_SRC_TABLE = 3,       /* S_v(x_i) tabulated by fvm_set_source_table                        */
    FVM_SRC_GRAY_SCOTT = 4,  /* neq=2: b(1-u) - u v^2 ; -d v + u v^2        params: (b, d)        */
    FVM_SRC_BRUSSELATOR = 5, /* neq=2: u^2 v - (b+1)... as the tutorial: u^2 v - 2u ; -u^2 v + u  */
    FVM_SRC_KELLER_SEGEL = 6 /* neq=2: u(1-u) ; u - a v                     params: a             */
};

/* Boundary / internal condition functions a(x,y,t,u,p): src/conditions.jl:16-20 */
enum fvm_cond_fn {
    FVM_COND_CONST = 0,    /* c0                                   */
    FVM_COND_AFFINE_U = 1, /* c0 + c1 * u_var                      */
    FVM_COND_EXP_SAT = 2,  /* c0 * (1 - exp(-t / c1))              */
    FVM_COND_LINEAR_XY = 3, /* c0 + c1 x + c2 y                    */
    FVM_COND_EXP_XYT = 4    /* c0 * exp(c1 x + c2 y + c3 t)        */
};

/* Linear templates: the constructors under src/specific_problems */
enum fvm_template {
    FVM_TPL_DIFFUSION = 0,                 /* diffusion_equation.jl:69-101 */
    FVM_TPL_LINEAR_REACTION_DIFFUSION = 1, /* linear_reaction_diffusion_equations.jl:76-125 */
    FVM_TPL_MEAN_EXIT_TIME = 2,            /* mean_exit_time.jl:57-94 */
    FVM_TPL_POISSON = 3,                   /* poissons_equation.jl:58-88 */
    FVM_TPL_LAPLACE = 4                    /* laplaces_equation.jl:51-78 */
};

enum fvm_krylov_method { FVM_KRYLOV_PCG = 0, FVM_KRYLOV_BICGSTAB = 1 };

/* ---- lifecycle ------------------------------------------------------------------ */

/* Replaces FVMGeometry(tri) (src/geometry.jl:99-169): takes the triangulation's points
 * (interleaved x,y) and solid triangles (ccw, stored rotation kept), `index_base` 1 for
 * Julia.  neq = 1 for FVMProblem, N for FVMSystem{N} (src/problem.jl:233-279). */
int32_t fvm_create(const double* xy, int64_t n_points, const int32_t* triangles, int64_t n_triangles,
                   int32_t index_base, int32_t neq, int32_t device, fvm_handle* out);

/* keys(get_boundary_edge_map(tri)) as directed ccw edges (u,v)
 * (src/equations/boundary_edge_contributions.jl:89-94). */
int32_t fvm_set_boundary_edges(fvm_handle h, const int32_t* uv, int64_t n_edges);

/* Conditions flattened per species (src/conditions.jl:506-544, src/problem.jl:322-357):
 * kind/fidx per boundary edge, in the order given to fvm_set_boundary_edges ... */
int32_t fvm_set_edge_conditions(fvm_handle h, int32_t var, const uint8_t* kind, const int32_t* fidx);
/* ... and per node (Dirichlet takes precedence over Dudt: resolve before calling,
 * src/equations/source_contributions.jl:5-12). */
int32_t fvm_set_node_conditions(fvm_handle h, int32_t var, const uint8_t* kind, const int32_t* fidx);
/* condition function `fidx` of species `var` (0-based) from the registry. */
int32_t fvm_set_condition_fn(fvm_handle h, int32_t var, int32_t fidx, int32_t fn_id, const double* params,
                             int32_t nparams);

int32_t fvm_set_flux(fvm_handle h, int32_t model, const double* params, int32_t nparams);
/* D at the 3T cv-edge midpoints (layout [T][3], caller triangle order) and at the 2 quarter
 * points of every boundary edge (layout [Eb][2]); for (x,y)-only diffusion functions. */
int32_t fvm_set_flux_table(fvm_handle h, const double* d_cv_edge, const double* d_bnd);
int32_t fvm_set_source(fvm_handle h, int32_t model, const double* params, int32_t nparams);
int32_t fvm_set_source_table(fvm_handle h, const double* s_node /* [N][neq] */);

/* Freezes the mesh: Hilbert-sorts triangles into tiles, renumbers nodes tile-major, builds the
 * per-tile gather lists that make the scatter deterministic (no fp64 atomics), computes the
 * geometry SoA on the device.  Options: tile_triangles (0 = measured default per kernel),
 * geometry_mode 0 = stored SoA (north_star layout), 1 = recomputed from vertex coordinates. */
int32_t fvm_finalize(fvm_handle h, int32_t tile_triangles, int32_t geometry_mode);
/* Host-only self-check of the planning fvm_finalize does before its first device call (Hilbert tiling, tile-major
 * renumbering, tile-local indices, gather lists, interface / partial-slot bookkeeping, live boundary edges, and the
 * tile packs of the streaming recompute kernel incl. their bank-conflict-free slot colouring): runs it
 * on the given mesh without a CUDA device and verifies the invariants the kernels rely on.  node_kind: [neq][N]
 * fvm_node_kind values or NULL (all free).  stats[8]: n_tiles, n_vertices, n_interface, n_partial, n_external,
 * max_local_nodes, n_live_boundary_edges, gather-list entries (= 3 T). */
int32_t fvm_plan_selftest(const double* xy, int64_t n_points, const int32_t* triangles, int64_t n_triangles,
                          int32_t index_base, int32_t neq, const int32_t* boundary_edges, int64_t n_edges,
                          const uint8_t* node_kind, int32_t tile_triangles, int64_t* stats);
int32_t fvm_destroy(fvm_handle h);
const char* fvm_last_error(fvm_handle h);
const char* fvm_version(void);

/* ---- the right-hand side (src/equations/main_equations.jl:28-44) -------------------- */

/* du = fvm_eqs!(du, u, p, t): triangle pass, boundary-edge pass, node pass. */
int32_t fvm_rhs(fvm_handle h, double t, const double* u, double* du, int32_t on_device);
/* same on device vectors already in native order (see fvm_to_native) */
int32_t fvm_rhs_native(fvm_handle h, double t, const double* u_native, double* du_native);
/* update_dirichlet_nodes! (src/equations/dirichlet.jl:78-86): u[i] = a(x_i,y_i,t,u[i]) */
int32_t fvm_apply_dirichlet(fvm_handle h, double t, double* u, int32_t on_device);
int32_t fvm_apply_dirichlet_native(fvm_handle h, double t, double* u_native);
int32_t fvm_to_native(fvm_handle h, const double* v_caller_dev, double* v_native_dev);
int32_t fvm_from_native(fvm_handle h, const double* v_native_dev, double* v_caller_dev);
/* Host vectors of >= 2^20 nodes take a banded pipeline: the caller-order vector is cut into bands, each band is
 * copied in and scattered to native order on a copy stream, the tiles (and boundary edges, interface nodes)
 * whose inputs have arrived run on the compute stream, and finished bands are gathered and copied out on a
 * third stream while later bands are still arriving (PCIe is full duplex).  Bit-identical to the plain
 * schedule.  Page-lock the buffers once with fvm_host_register to get the full rate (a Julia Vector or NumPy
 * array is pageable); unregister before freeing them.  Sharded handles do their halo
 * exchange at the head of the last stage.  A stage sends back the longest prefix of the output that is final after
 * it (not whole input bands).  Environment: FVM_NO_PIPELINE=1, FVM_PIPE_BANDS=K, FVM_PIPE_TAPER=0|1|2 (uniform bands /
 * first and last two at half width / quarter, quarter, half), FVM_PIPE_TRACE=1 (per-band device timeline on stderr),
 * FVM_PIPE_ZC=1|2|3 (page-locked buffers only: copy-in / copy-out / both through zero-copy kernels instead of the copy
 * engines; measured slower on B200, off by default). */
int32_t fvm_host_register(void* host_ptr, int64_t nbytes);
int32_t fvm_host_unregister(void* host_ptr);
/* the *_native calls are asynchronous on the handle's stream */
int32_t fvm_stream_synchronize(fvm_handle h);
int32_t fvm_get_stream(fvm_handle h, void** cuda_stream);
/* CUDA-event timing of the dominant kernel (RHS tile kernel / SpMV) on the launching stream:
 * arm with the number of launches to record, read back the summed duration. */
int32_t fvm_set_profiling(fvm_handle h, int32_t max_launches);
int32_t fvm_get_profile(fvm_handle h, double* total_ms, int64_t* launches);

/* ---- geometry / connectivity read-back for parity (src/geometry.jl:21-49) ----------- */
/* caller order; any pointer may be NULL.  V[N]; s9[T][9]; mid6[T][3][2]; nrm6[T][3][2]; len3[T][3] */
int32_t fvm_get_geometry(fvm_handle h, double* V, double* s9, double* mid6, double* nrm6, double* len3);
/* Self-check of geometry_mode 1: counts the triangles on which the division-free recomputed geometry of the
 * streaming kernel differs from the individually rounded reference arithmetic (src/geometry.jl:107-161): any bit at all
 * in the variant the u-dependent flux models run; any bit of s7..s9, the cv-edge midpoints or the cv-edge vectors, or
 * more than one ulp of s1..s6, in the variant of the fluxes that read alpha and beta only.  0 on every mesh tested. */
int32_t fvm_check_recompute_geometry(fvm_handle h, int64_t* n_mismatch);
/* native permutations: node_perm[new] = old (N), tri_perm[new] = old (T); tile layout stats */
int32_t fvm_get_permutation(fvm_handle h, int32_t* node_perm, int32_t* tri_perm);
int32_t fvm_get_stats(fvm_handle h, int64_t* stats /* [16] */);

/* ---- linear templates (src/specific_problems/abstract_templates.jl:73-325) ---------- */

/* Assembles A (CSR on the structural pattern of jacobian_sparsity, src/solve.jl:56-77) and b on
 * the device.  d_const is used when d_cv_edge == NULL.  node_value[N]: value of the node's
 * condition function (Dirichlet value for u0 / steady b, Dudt value for b); edge_value[Eb][2]:
 * Neumann function at the quarter points; source[N]: f(x_i) for Poisson, the diagonal term for
 * linear reaction-diffusion; pointers may be NULL where the template does not use them.
 * reference_quirks != 0 reproduces abstract_templates.jl:262 (j row scaled by 1/V_i). */
int32_t fvm_assemble(fvm_handle h, int32_t template_id, double d_const, const double* d_cv_edge,
                     const double* d_bnd, const double* node_value, const double* edge_value,
                     const double* source, int32_t reference_quirks);
int32_t fvm_get_csr_size(fvm_handle h, int64_t* n_rows, int64_t* nnz);
/* caller numbering, columns sorted within a row, explicit zeros kept on the structural pattern */
int32_t fvm_get_csr(fvm_handle h, int32_t* rowptr, int32_t* col, double* val, double* b);
/* y = A x + b (add_b != 0) or y = A x : the MatrixOperator mul! of diffusion_equation.jl:93-94 */
int32_t fvm_spmv(fvm_handle h, const double* x, double* y, int32_t add_b, int32_t on_device);
int32_t fvm_spmv_native(fvm_handle h, const double* x_native, double* y_native, int32_t add_b);

/* Device-resident fixed-step Tsit5 (solve(prob, Tsit5(); adaptive=false, dt)).  use_operator != 0
 * integrates du/dt = A u + b (templates), else du/dt = fvm_eqs!(u,t) with the Dirichlet callback
 * after every step.  u: in u0, out u(t1), caller order, host or device.  usave[nsave][len] receives
 * the states at tsave (each must be a step boundary). */
int32_t fvm_tsit5(fvm_handle h, int32_t use_operator, double* u, double t0, double t1, double dt,
                  int64_t nsave, const double* tsave, double* usave, int32_t on_device);

/* Adaptive Tsit5 (solve(prob, Tsit5(); abstol, reltol, saveat)): embedded error estimate, PI step
 * controller with OrdinaryDiffEq's constants, saveat times hit exactly.  dt0 <= 0 selects the initial
 * step automatically.  Returns the accepted / rejected step counts. */
int32_t fvm_tsit5_adaptive(fvm_handle h, int32_t use_operator, double* u, double t0, double t1, double abstol,
                           double reltol, double dt0, int64_t nsave, const double* tsave, double* usave,
                           int32_t on_device, int64_t* n_accept, int64_t* n_reject);

/* Steady templates: solves A x = b with Jacobi-preconditioned CG (on the symmetrised system
 * -V A, Dirichlet-consistent start) or BiCGStab.  x: in initial guess, out solution. */
int32_t fvm_krylov(fvm_handle h, int32_t method, double* x, double rtol, int32_t maxit, int32_t* iters,
                   double* relres, int32_t on_device);

/* ---- sparse Jacobian of fvm_eqs! (jacobian_sparsity, src/solve.jl:56-131) ------------------------ */
/* J = d(du)/du at (u, t), exact derivatives of the registered flux / source / condition functions
 * (forward-mode duals on the device); pattern = jacobian_sparsity, systems node-major interleaved. */
int32_t fvm_jacobian(fvm_handle h, double t, const double* u, int32_t on_device);
int32_t fvm_get_jacobian_size(fvm_handle h, int64_t* n_rows, int64_t* nnz);
/* caller numbering; val may be NULL to fetch the pattern only (the jac_prototype of solve.jl:170) */
int32_t fvm_get_jacobian_csr(fvm_handle h, int32_t* rowptr, int32_t* col, double* val);

/* solve(SteadyFVMProblem(prob), NewtonRaphson()) (src/solve.jl:209-220) on the device: u <- u - J(u)^-1 F(u) with
 * F = fvm_eqs!(u, t), J = fvm_jacobian and every Newton system solved by Jacobi-preconditioned BiCGStab on the block CSR
 * (rows of J without a non-zero -- Dirichlet nodes, points that are not vertices -- keep their value).  Stops when
 * max|F| <= abstol + reltol * max|F(u0)| or after maxiters iterations.  u: in the initial guess, out the iterate
 * (caller order, host or device; Dirichlet nodes are set to their condition values first).  Outputs (each may be NULL):
 * Newton iterations taken, final and initial max|F|, total BiCGStab iterations. */
int32_t fvm_newton(fvm_handle h, double t, double* u, double abstol, double reltol, int32_t maxiters, double lin_rtol,
                   int32_t lin_maxit, int32_t* iters, double* resid, double* resid0, int64_t* lin_iters, int32_t on_device);

/* ---- post-processing (src/utils.jl:23-27 pl_interpolate, src/problem.jl:458-487 compute_flux) ---- */
/* For n query points (x,y) lying in given triangles (caller triangle indices): the piecewise-linear
 * interpolant alpha*x + beta*y + gamma of u (nrm == NULL), or q(x,y,t,alpha,beta,gamma) . nrm.
 * tri_idx, xy, nrm, out are host buffers; out is [n][neq]. */
int32_t fvm_eval_points(fvm_handle h, double t, const double* u, int32_t u_on_device, int64_t n, const int32_t* tri_idx,
                        const double* xy, const double* nrm, double* out);

/* ---- multi-GPU: node partition + one-layer ghost halo per rank ---------------- */
/* Each rank creates its handle on the LOCAL mesh: owned nodes + ghost nodes, and every triangle that
 * touches an owned node.  Call order: fvm_create, setters, fvm_set_ghost_nodes, fvm_finalize,
 * fvm_shard_init, fvm_set_halo.  Afterwards fvm_rhs*, fvm_spmv* and fvm_tsit5 refresh the ghost
 * entries of their input vector with grouped ncclSend/ncclRecv before computing; outputs are valid
 * on owned nodes (ghost rows are 0). */
/* METIS-style node partition (north_star d): multilevel recursive bisection of the node graph (an edge per
 * pair of nodes sharing a triangle, the pattern of jacobian_sparsity, src/solve.jl:56-77): heavy-edge matching,
 * greedy graph growing on the coarsest graph, Fiduccia-Mattheyses refinement while uncoarsening; parts differ
 * by at most one node.  Deterministic.  Host-only. */
int32_t fvm_partition_graph(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                            int32_t n_parts, int32_t* owner /* [n_points], 0-based part of every node */);
int32_t fvm_partition_edge_cut(int64_t n_points, const int32_t* triangles, int64_t n_triangles, int32_t index_base,
                               const int32_t* owner, int64_t* cut);
int32_t fvm_set_ghost_nodes(fvm_handle h, const uint8_t* is_ghost /* [N], 1 = owned by another rank */);
/* nccl_unique_id: the 128-byte ncclUniqueId created on rank 0 and broadcast by the host. */
int32_t fvm_nccl_unique_id(void* out128);
int32_t fvm_shard_init(fvm_handle h, const void* nccl_unique_id, int32_t rank, int32_t nranks);
/* neighbours and the (local) node lists to send to / receive from each; both sides of a pair must
 * list the shared nodes in the same order (ascending global node id). */
int32_t fvm_set_halo(fvm_handle h, int32_t n_neighbours, const int32_t* neighbour_ranks, const int32_t* send_ptr,
                     const int32_t* send_nodes, const int32_t* recv_ptr, const int32_t* recv_nodes);
int32_t fvm_halo_exchange_native(fvm_handle h, double* u_native);
/* How the ghost refresh runs: 0 none (single rank / no neighbours), 1 grouped ncclSend/ncclRecv, 2 peer-mapped: the
 * neighbours' receive slabs are mapped through CUDA IPC and ONE kernel stores the owned boundary values straight into
 * them over NVLink (epoch flag with system-scope release), ONE kernel waits for the neighbours' epochs and scatters.
 * Chosen at fvm_set_halo (unanimously over the ranks; FVM_HALO_PEER=0 forces NCCL).  *timed_out != 0: a neighbour
 * never signalled within ~10 s and the waiting kernel gave up. */
int32_t fvm_halo_mode(fvm_handle h, int32_t* mode, int32_t* timed_out);

/* ---- FVMWIRE: flat binary SoA container for meshes and solutions ------------
 * The reference has no file format: a mesh is the DelaunayTriangulation object FVMGeometry(tri) walks
 * (src/geometry.jl:99-106) and a solution is sol.u / sol.t (src/solve.jl:197-208).  The container holds
 * exactly those arrays, column-major (dims[0] fastest, i.e. Julia's size(A)), 64-byte aligned so a payload
 * can be handed to cudaMemcpy / mmap as it lies: 64-byte header, a fixed table of FVM_WIRE_MAX_ARRAYS
 * 96-byte entries {name[32], dtype, rank, dims[4], offset, nbytes, crc32}, then the payloads.  Every
 * payload and the table carry a CRC-32 (zlib polynomial); fvm_wire_open / fvm_wire_get fail with
 * FVM_ERR_IO on truncation or corruption.  Host-only calls: they need no CUDA device.
 * Mesh schema: "points" f64 (2,N); "triangles" i32 (3,T); "index_base" i32 (1); "boundary_ptr" i32 (S+1)
 * + "boundary_nodes" i32 (ccw node sequence of every boundary section, get_boundary_nodes(tri));
 * optional "boundary_edges" i32 (2,Eb).  Solution schema: "t" f64 (nsave); "u" f64 (N,nsave) or
 * (neq,N,nsave). */
#define FVM_WIRE_MAX_ARRAYS 64
enum fvm_wire_dtype { FVM_WIRE_F64 = 1, FVM_WIRE_I32 = 2, FVM_WIRE_U8 = 3, FVM_WIRE_I64 = 4 };
typedef struct fvm_wire* fvm_wire_handle;

int32_t fvm_wire_create(const char* path, fvm_wire_handle* out); /* write mode; arrays are streamed to disk */
int32_t fvm_wire_put(fvm_wire_handle w, const char* name, int32_t dtype, int32_t rank, const int64_t* dims,
                     const void* data);
int32_t fvm_wire_open(const char* path, fvm_wire_handle* out);   /* read mode; validates header + table */
int32_t fvm_wire_count(fvm_wire_handle w, int32_t* n_arrays);
int32_t fvm_wire_info(fvm_wire_handle w, int32_t index, char* name32, int32_t* dtype, int32_t* rank,
                      int64_t* dims4, int64_t* nbytes);
int32_t fvm_wire_find(fvm_wire_handle w, const char* name, int32_t* index);
int32_t fvm_wire_get(fvm_wire_handle w, int32_t index, void* out, int64_t nbytes); /* verifies the CRC */
int32_t fvm_wire_close(fvm_wire_handle w);                       /* write mode: finishes header + table */
const char* fvm_wire_last_error(fvm_wire_handle w);              /* w == NULL: last failing open/create */
uint32_t fvm_wire_crc32(const void* data, int64_t nbytes);
/* fvm_create + fvm_set_boundary_edges from a mesh container (replaces FVMGeometry(tri) for meshes that
 * are produced elsewhere); continue with the setters and fvm_finalize. */
int32_t fvm_create_from_wire(const char* path, int32_t neq, int32_t device, fvm_handle* out);

#ifdef __cplusplus
}
#endif
#endif /* FVMCUDA_H */
