"""FVMGeometry / FVMProblem / FVMSystem / SteadyFVMProblem and the GPU `fvm_eqs!`.

Mirrors /root/reference/src/geometry.jl, src/problem.jl and src/solve.jl:1-42: the constructors
keep the reference's names, keyword arguments and assertions; `get_cuda_parameters(prob)` is the
sibling of `get_multithreading_parameters` that holds the libfvmcuda handle, and
`fvm_eqs(du, u, p, t)` is `fvm_eqs!` for `p.parallel == "cuda"`."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import functors as F
from .conditions import BoundaryConditions, Conditions, InternalConditions


class FVMGeometry:
    """geometry.jl:44-49.  The control-volume data lives on the device; `cv_volumes` and
    `triangle_props` are read back on demand (parity checks, host-side tabulation)."""

    def __init__(self, tri):
        self.triangulation = tri
        self._geom = None

    def __repr__(self):  # Base.show, geometry.jl:50-55
        tri = self.triangulation
        t = tri.triangles.astype(np.int64)
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
        ne = len(np.unique(e[:, 0] * tri.num_points + e[:, 1]))
        return "FVMGeometry with %d control volumes, %d triangles, and %d edges" % (int(tri.solid_vertex_mask().sum()), tri.num_triangles, ne)

    def _fetch(self):
        if self._geom is None:
            tri = self.triangulation
            h = _create_handle(tri, 1)
            try:
                L.check(h, L.lib().fvm_finalize(h, 0, 0))
                N, T = tri.num_points, tri.num_triangles
                V = np.empty(N)
                s9 = np.empty((T, 9))
                mid = np.empty((T, 3, 2))
                nrm = np.empty((T, 3, 2))
                ln = np.empty((T, 3))
                L.check(h, L.lib().fvm_get_geometry(h, L.dp(V), L.dp(s9), L.dp(mid), L.dp(nrm), L.dp(ln)))
                self._geom = dict(cv_volumes=V, s=s9, mid=mid, nrm=nrm, len=ln)
            finally:
                L.lib().fvm_destroy(h)
        return self._geom

    @property
    def cv_volumes(self):
        return self._fetch()["cv_volumes"]

    @property
    def triangle_props(self):
        """dict of arrays in the caller's triangle order: s (T,9), mid (T,3,2), nrm (T,3,2), len (T,3)"""
        return self._fetch()

    # host-side evaluation points for (x,y)-only coefficient tabulation, geometry.jl:114-151
    def cv_edge_midpoints(self):
        P, Tr = self.triangulation.points, self.triangulation.triangles
        p, q, r = P[Tr[:, 0]], P[Tr[:, 1]], P[Tr[:, 2]]
        c = (p + q + r) / 3
        out = np.empty((len(Tr), 3, 2))
        for e, (a, b) in enumerate(((p, q), (q, r), (r, p))):
            out[:, e, :] = ((a + b) / 2 + c) / 2
        return out

    def boundary_quarter_points(self, uv):
        """control_volumes.jl:41-56: m_i, m_j of every boundary edge, shape (Eb,2,2)."""
        P = self.triangulation.points
        p, q = P[uv[:, 0]], P[uv[:, 1]]
        mij = (p + q) / 2
        return np.stack([(p + mij) / 2, (q + mij) / 2], axis=1)


def _create_handle(tri, neq, device=None, mesh_file=None):
    if device is None:
        device = _current_device()
    h = L.H()
    if mesh_file is not None:  # points / triangles come from an FVMWIRE container (wire.py)
        rc = L.lib().fvm_create_from_wire(str(mesh_file).encode(), neq, device, C.byref(h))
        if rc != L.OK:
            msg = L.lib().fvm_wire_last_error(None) if rc == L.ERR_IO else L.lib().fvm_last_error(None)
            raise L.FVMCudaError(rc, msg.decode())
        return h
    pts = L.f64(tri.points)
    tr = L.i32(tri.triangles)
    rc = L.lib().fvm_create(L.dp(pts), tri.num_points, L.ip(tr), tri.num_triangles, 0, neq, device, C.byref(h))
    if rc != L.OK:
        raise L.FVMCudaError(rc, L.lib().fvm_last_error(None).decode())
    return h


def _current_device():
    import os
    return int(os.environ.get("FVM_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def construct_flux_function(q, D, Dp=None):
    """problem.jl:425-440: a diffusion function is turned into a flux spec."""
    if q is not None:
        return q
    if D is None:
        raise AssertionError("either flux_function or diffusion_function must be given")
    if isinstance(D, (int, float)):
        return F.ConstantDiffusion(float(D))
    return D


_FLUX_TYPES = (F.ConstantDiffusion, F.TabulatedDiffusion, F.PowerDiffusion, F.AdvectionDiffusionFlux, F.KellerSegelFlux)
_SRC_TYPES = (F.ZeroSource, F.LinearSource, F.LogisticSource, F.TabulatedSource, F.GrayScottSource, F.BrusselatorSource,
              F.KellerSegelSource)


FLUX_ERROR = """
The flux function errored when evaluated. Please recheck your specification of the function.
If any of your problems have been defined in terms of a diffusion function D(x, y, t, u, p)
rather than a flux function, then you need to instead provide a flux function q(x, y, t, α, β, γ, p),
recalling the relationship between the two:

    q(x, y, t, α, β, γ, p) = -D(x, y, t, α*x + β*y + γ, p) .* (α, β)

where p is the same argument for both functions.
"""


class InvalidFluxError(Exception):
    """problem.jl:291-316: a member of an FVMSystem was defined through `diffusion_function`, whose scalar
    flux cannot take the tuples (α, β, γ) of all species (test/equations.jl:108)."""

    def __init__(self):
        super().__init__(FLUX_ERROR)


class FVMProblem:
    """problem.jl:96-163."""

    def __init__(self, mesh, boundary_conditions, internal_conditions=None, *,
                 diffusion_function=None, diffusion_parameters=None,
                 source_function=None, source_parameters=None,
                 flux_function=None, flux_parameters=None,
                 initial_condition, initial_time=0.0, final_time):
        ic = np.ascontiguousarray(initial_condition, dtype=np.float64)
        if ic.ndim != 1 or len(ic) != mesh.triangulation.num_points:
            raise AssertionError("The initial condition must have the same number of elements as the number of nodes "
                                 "in the mesh (including nodes that aren't vertices in the mesh itself).")
        self.mesh = mesh
        if isinstance(boundary_conditions, Conditions):
            self.conditions = boundary_conditions
        else:
            self.conditions = Conditions(mesh, boundary_conditions, internal_conditions or InternalConditions())
        self.flux_function = construct_flux_function(flux_function, diffusion_function, diffusion_parameters)
        self._flux_from_diffusion = flux_function is None  # problem.jl:425-440 wraps D into a SCALAR flux function
        self.flux_parameters = flux_parameters
        self.source_function = source_function if source_function is not None else F.ZeroSource()
        self.source_parameters = source_parameters
        self.initial_condition = ic
        self.initial_time = float(initial_time)
        self.final_time = float(final_time)
        _check_registry(self.flux_function, _FLUX_TYPES, "flux/diffusion")
        _check_registry(self.source_function, _SRC_TYPES, "source")

    neqs = 0
    problems = property(lambda self: (self,))

    def __repr__(self):
        nv = int(self.mesh.triangulation.solid_vertex_mask().sum())
        return "FVMProblem with %d nodes and time span (%s, %s)" % (nv, self.initial_time, self.final_time)


def _check_registry(fn, types, what):
    if not isinstance(fn, types):
        raise L.UnsupportedClosureError(
            L.ERR_UNSUPPORTED,
            "%s function %r is not in the compiled device registry (%s); arbitrary closures cannot run on the GPU. "
            "(x,y)-only functions can be wrapped in TabulatedDiffusion / TabulatedSource."
            % (what, fn, ", ".join(t.__name__ for t in types)))


class FVMSystem:
    """problem.jl:233-279, 359-411: state is (N, neq) C-order == Julia's Matrix(neq, N)."""

    def __init__(self, *probs):
        if len(probs) == 0:
            raise AssertionError("There must be at least one problem.")
        self.problems = tuple(probs)
        self.mesh = probs[0].mesh
        if not all(p.mesh is self.mesh for p in probs):
            raise AssertionError("All problems must have the same mesh.")
        if not all(p.initial_time == probs[0].initial_time for p in probs):
            raise AssertionError("All problems must have the same initial time.")
        if not all(p.final_time == probs[0].final_time for p in probs):
            raise AssertionError("All problems must have the same final time.")
        self.initial_time, self.final_time = probs[0].initial_time, probs[0].final_time
        self.neqs = len(probs)
        self.initial_condition = np.ascontiguousarray(np.stack([p.initial_condition for p in probs], axis=1))
        self.conditions = tuple(p.conditions for p in probs)
        nf = [len(p.conditions.functions) for p in probs]
        self.cnum_fncs = tuple(int(x) for x in np.concatenate([[0], np.cumsum(nf)[:-1]]))  # problem.jl:380-383
        if any(getattr(p, "_flux_from_diffusion", False) for p in probs):  # _check_fvmsystem_flux_function, problem.jl:295-316
            raise InvalidFluxError()

    def __repr__(self):
        return "FVMSystem with %d equations and time span (%s, %s)" % (self.neqs, self.initial_time, self.final_time)


class SteadyFVMProblem:
    """problem.jl:174-180."""

    def __init__(self, prob):
        self.problem = prob
        self.neqs = prob.neqs

    def __repr__(self):  # Base.show, problem.jl:182-190
        nv = int(self.problem.mesh.triangulation.solid_vertex_mask().sum())
        if self.neqs:
            return "SteadyFVMProblem with %d nodes and %d equations" % (nv, self.neqs)
        return "SteadyFVMProblem with %d nodes" % nv


# ---------------------------------------------------------------------------------------------
class Engine:
    """Owns one libfvmcuda handle for a problem (RHS) or a template (linear operator)."""

    def __init__(self, mesh, neq, conditions, flux=None, source=None, tile_triangles=0, geometry_mode=1, device=None, ghost=None,
                 mesh_file=None):
        lib = L.lib()
        tri = mesh.triangulation
        self.mesh, self.neq = mesh, neq
        self.N, self.T = tri.num_points, tri.num_triangles
        self.h = _create_handle(tri, neq, device, mesh_file)
        self._keep = []
        self.validation_error = None  # sharded templates: rank-local verdict, combined over ranks by install_halo
        try:
            uv = conditions[0].boundary_edges
            self.boundary_edges = uv
            base = 0
            if mesh_file is not None:  # the handle indexes like the file does (1-based when Julia wrote it)
                from .wire import WireReader
                with WireReader(mesh_file) as r:
                    base = int(r.get("index_base")[0]) if "index_base" in r else 1
            L.check(self.h, lib.fvm_set_boundary_edges(self.h, L.ip(L.i32(uv + base)), len(uv)))
            for v, c in enumerate(conditions):
                ek, ef = np.ascontiguousarray(c.edge_kind, np.uint8), L.i32(c.edge_fidx)
                nk, nf = np.ascontiguousarray(c.node_kind, np.uint8), L.i32(c.node_fidx)
                L.check(self.h, lib.fvm_set_edge_conditions(self.h, v, L.bp(ek), L.ip(ef)))
                L.check(self.h, lib.fvm_set_node_conditions(self.h, v, L.bp(nk), L.ip(nf)))
                used = set(np.unique(nf[nk != 0]).tolist()) | set(np.unique(ef[ek == 1]).tolist())
                for fidx in sorted(used):
                    fid, params = F.cond_spec(c.functions[fidx])
                    p = L.f64(params)
                    L.check(self.h, lib.fvm_set_condition_fn(self.h, v, fidx, fid, L.dp(p), len(p)))
            if flux is not None:
                self._set_flux(flux, uv)
            if source is not None:
                self._set_source(source)
            if ghost is not None:
                gh = np.ascontiguousarray(ghost, dtype=np.uint8)
                L.check(self.h, lib.fvm_set_ghost_nodes(self.h, L.bp(gh)))
            L.check(self.h, lib.fvm_finalize(self.h, tile_triangles, geometry_mode))
        except Exception:
            lib.fvm_destroy(self.h)
            self.h = None
            raise

    def _set_flux(self, specs, uv):
        lib = L.lib()
        s0 = specs[0]
        if not all(type(s) is type(s0) for s in specs):
            raise L.UnsupportedClosureError(L.ERR_UNSUPPORTED, "all species of an FVMSystem must use the same flux model class")
        if isinstance(s0, F.ConstantDiffusion):
            model, p = F.FLUX_DIFF_CONST, [s.D for s in specs]
        elif isinstance(s0, F.TabulatedDiffusion):
            if not all(s.fn is s0.fn for s in specs):
                raise L.UnsupportedClosureError(L.ERR_UNSUPPORTED, "a tabulated diffusion function must be shared by all species")
            mid = self.mesh.cv_edge_midpoints()
            dt = L.f64(np.broadcast_to(s0.fn(mid[..., 0], mid[..., 1]), mid.shape[:2]))
            qp = self.mesh.boundary_quarter_points(uv) if len(uv) else np.zeros((0, 2, 2))
            db = L.f64(np.broadcast_to(s0.fn(qp[..., 0], qp[..., 1]), qp.shape[:2]))
            L.check(self.h, lib.fvm_set_flux_table(self.h, L.dp(dt), L.dp(db)))
            return
        elif isinstance(s0, F.PowerDiffusion):
            model, p = F.FLUX_DIFF_POWER, [x for s in specs for x in (s.D0, s.m, 1.0 if s.use_abs else 0.0)]
        elif isinstance(s0, F.AdvectionDiffusionFlux):
            model, p = F.FLUX_ADVDIFF, [x for s in specs for x in (s.D, s.nu_x, s.nu_y)]
        elif isinstance(s0, F.KellerSegelFlux):
            if not all(s == s0 for s in specs):
                raise AssertionError("KellerSegelFlux must be the same spec on every species")
            model, p = F.FLUX_KELLER_SEGEL, [s0.c, s0.D]
        else:
            _check_registry(s0, _FLUX_TYPES, "flux/diffusion")
        p = L.f64(p)
        L.check(self.h, lib.fvm_set_flux(self.h, model, L.dp(p), len(p)))

    def _set_source(self, specs):
        lib = L.lib()
        s0 = specs[0]
        system_specs = (F.GrayScottSource, F.BrusselatorSource, F.KellerSegelSource)
        if isinstance(s0, system_specs):
            if not all(s == s0 for s in specs):
                raise AssertionError("%s must be the same spec on every species" % type(s0).__name__)
            model = {F.GrayScottSource: F.SRC_GRAY_SCOTT, F.BrusselatorSource: F.SRC_BRUSSELATOR,
                     F.KellerSegelSource: F.SRC_KELLER_SEGEL}[type(s0)]
            p = {F.GrayScottSource: lambda s: [s.b, s.d], F.BrusselatorSource: lambda s: [],
                 F.KellerSegelSource: lambda s: [s.a]}[type(s0)](s0)
        elif all(isinstance(s, F.ZeroSource) for s in specs):
            return
        elif any(isinstance(s, F.TabulatedSource) for s in specs):
            P = self.mesh.triangulation.points
            tab = np.zeros((self.N, self.neq))
            for v, s in enumerate(specs):
                if isinstance(s, F.TabulatedSource):
                    tab[:, v] = s.fn(P[:, 0], P[:, 1])
                elif not isinstance(s, F.ZeroSource):
                    raise L.UnsupportedClosureError(L.ERR_UNSUPPORTED, "tabulated sources can only be mixed with ZeroSource")
            tab = L.f64(tab)
            L.check(self.h, lib.fvm_set_source_table(self.h, L.dp(tab)))
            return
        elif all(isinstance(s, (F.LinearSource, F.ZeroSource)) for s in specs):
            model = F.SRC_LINEAR
            p = [x for s in specs for x in ((s.lam, s.mu) if isinstance(s, F.LinearSource) else (0.0, 0.0))]
        elif all(isinstance(s, (F.LogisticSource, F.ZeroSource)) for s in specs):
            model = F.SRC_LOGISTIC
            p = [s.lam if isinstance(s, F.LogisticSource) else 0.0 for s in specs]
        else:
            raise L.UnsupportedClosureError(L.ERR_UNSUPPORTED, "this combination of per-species source models is not compiled")
        p = L.f64(p)
        L.check(self.h, lib.fvm_set_source(self.h, model, L.dp(p) if len(p) else None, len(p)))

    # ---- calls ----
    def rhs(self, du, u, t):
        L.check(self.h, L.lib().fvm_rhs(self.h, float(t), u.ctypes.data, du.ctypes.data, 0))
        return du

    def rhs_device(self, du_ptr, u_ptr, t, native=False):
        if native:
            L.check(self.h, L.lib().fvm_rhs_native(self.h, float(t), u_ptr, du_ptr))
        else:
            L.check(self.h, L.lib().fvm_rhs(self.h, float(t), u_ptr, du_ptr, 1))

    def spmv_device(self, y_ptr, x_ptr, add_b=True, native=False):
        if native:
            L.check(self.h, L.lib().fvm_spmv_native(self.h, x_ptr, y_ptr, 1 if add_b else 0))
        else:
            L.check(self.h, L.lib().fvm_spmv(self.h, x_ptr, y_ptr, 1 if add_b else 0, 1))

    def tsit5_device(self, u_ptr, t0, t1, dt, use_operator):
        """device-resident fixed-step Tsit5 on a device vector in caller order"""
        L.check(self.h, L.lib().fvm_tsit5(self.h, 1 if use_operator else 0, u_ptr, float(t0), float(t1), float(dt), 0, None, None, 1))

    def apply_dirichlet(self, u, t):
        L.check(self.h, L.lib().fvm_apply_dirichlet(self.h, float(t), u.ctypes.data, 0))
        return u

    def synchronize(self):
        L.check(self.h, L.lib().fvm_stream_synchronize(self.h))

    def stream(self):
        s = C.c_void_p()
        L.check(self.h, L.lib().fvm_get_stream(self.h, C.byref(s)))
        return s.value or 0

    def set_profiling(self, max_launches):
        L.check(self.h, L.lib().fvm_set_profiling(self.h, int(max_launches)))

    def get_profile(self):
        """(summed ms, launches) of the dominant kernel since profiling was armed"""
        ms = C.c_double()
        n = C.c_int64()
        L.check(self.h, L.lib().fvm_get_profile(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stats(self):
        st = np.zeros(16, dtype=np.int64)
        L.check(self.h, L.lib().fvm_get_stats(self.h, st.ctypes.data_as(L.c_lp)))
        keys = ["n_tiles", "tile_triangles", "n_vertices", "n_interface", "n_partial", "n_external", "max_local_nodes",
                "n_live_boundary_edges", "n_dirichlet", "smem_bytes", "nnz", "max_row", "pipe_bands", "pipe_early_bands", "pipe_calls"]
        out = dict(zip(keys, st.tolist()))
        names = ("undecided", "pipeline", "plain")
        out["host_schedule_rhs"], out["host_schedule_spmv"] = names[int(st[15]) & 3], names[(int(st[15]) >> 2) & 3]
        return out

    def halo_mode(self):
        """(mode, timed_out): mode 0 no halo, 1 grouped ncclSend/ncclRecv, 2 peer-mapped NVLink stores (fvm_halo_mode)."""
        import ctypes as C
        mode, to = C.c_int32(), C.c_int32()
        L.check(self.h, L.lib().fvm_halo_mode(self.h, C.byref(mode), C.byref(to)))
        return mode.value, to.value

    def permutation(self):
        node = np.empty(self.N, dtype=np.int32)
        tri = np.empty(self.T, dtype=np.int32)
        L.check(self.h, L.lib().fvm_get_permutation(self.h, L.ip(node), L.ip(tri)))
        return node, tri

    def geometry(self):
        V = np.empty(self.N)
        s9 = np.empty((self.T, 9))
        mid = np.empty((self.T, 3, 2))
        nrm = np.empty((self.T, 3, 2))
        ln = np.empty((self.T, 3))
        L.check(self.h, L.lib().fvm_get_geometry(self.h, L.dp(V), L.dp(s9), L.dp(mid), L.dp(nrm), L.dp(ln)))
        return dict(cv_volumes=V, s=s9, mid=mid, nrm=nrm, len=ln)

    def close(self):
        if getattr(self, "h", None):
            L.lib().fvm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaParameters:
    """The `p` of fvm_eqs!(du,u,p,t) (main_equations.jl:8-26) for the GPU path."""
    parallel = "cuda"

    def __init__(self, prob, engine):
        self.prob = prob
        self.engine = engine


def get_cuda_parameters(prob, tile_triangles=0, geometry_mode=1, device=None, ghost=None, mesh_file=None):
    """Sibling of get_multithreading_parameters (solve.jl:1-27): builds the device state once.
    `mesh_file`: let the library read points / triangles from an FVMWIRE container of the same mesh."""
    if isinstance(prob, SteadyFVMProblem):
        prob = prob.problem
    probs = prob.problems
    neq = max(1, prob.neqs)
    conds = [p.conditions for p in probs]
    eng = Engine(prob.mesh, neq, conds, [p.flux_function for p in probs], [p.source_function for p in probs],
                 tile_triangles, geometry_mode, device, ghost, mesh_file)
    return CudaParameters(prob, eng)


class pinned:
    """Context manager / helper that page-locks NumPy arrays for the host-buffer calls (fvm_host_register):

        with G.pinned(u, du):
            G.fvm_eqs(du, u, p, t)      # full PCIe rate, copy-in / kernels / copy-out overlapped
    """

    def __init__(self, *arrays):
        self.arrays = arrays
        for a in arrays:
            rc = L.lib().fvm_host_register(a.ctypes.data, a.nbytes)
            if rc != L.OK:
                raise L.FVMCudaError(rc, L.lib().fvm_last_error(None).decode())

    def release(self):
        for a in self.arrays:
            L.lib().fvm_host_unregister(a.ctypes.data)
        self.arrays = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()


def fvm_eqs(du, u, p, t):
    """fvm_eqs!(du, u, p, t) (main_equations.jl:28-35) for p.parallel == "cuda".  `u`/`du` are host
    float64 arrays of shape (N,) or (N, neq) in the caller's node order; mutates and returns du."""
    if u.dtype != np.float64 or du.dtype != np.float64:
        raise TypeError("the GPU path supports Float64 only (ForwardDiff duals are rejected)")
    if u.shape != p.prob.initial_condition.shape or du.shape != u.shape:
        raise AssertionError("u and du must have the shape of the initial condition")
    if not (u.flags.c_contiguous and du.flags.c_contiguous):
        raise AssertionError("u and du must be contiguous")
    return p.engine.rhs(du, u, t)


def update_dirichlet_nodes(u, t, p):
    """update_dirichlet_nodes!(integrator) (dirichlet.jl:78-86)."""
    return p.engine.apply_dirichlet(u, t)


def _jac_csr(p, with_values):
    import scipy.sparse as sp
    h = p.engine.h
    n, nnz = C.c_int64(), C.c_int64()
    L.check(h, L.lib().fvm_get_jacobian_size(h, C.byref(n), C.byref(nnz)))
    rowptr = np.empty(n.value + 1, np.int32)
    col = np.empty(nnz.value, np.int32)
    val = np.empty(nnz.value) if with_values else None
    L.check(h, L.lib().fvm_get_jacobian_csr(h, L.ip(rowptr), L.ip(col), L.dp(val)))
    if val is None:
        val = np.ones(nnz.value)
    return sp.csr_matrix((val, col, rowptr), shape=(n.value, n.value))


def jacobian_sparsity(p):
    """jacobian_sparsity(prob) (solve.jl:50-131): the prototype (ones on the structural pattern); for an
    FVMSystem rows/columns are node-major interleaved, (i-1)*neq + l in the reference's 1-based terms."""
    return _jac_csr(p, False)


def jacobian(u, p, t):
    """Sparse Jacobian of fvm_eqs! at (u, t) as a SciPy CSR matrix, assembled on the device with exact
    derivatives (what the reference obtains by pushing ForwardDiff duals through fvm_eqs!, solve.jl:5)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    L.check(p.engine.h, L.lib().fvm_jacobian(p.engine.h, float(t), u.ctypes.data, 0))
    return _jac_csr(p, True)


class _EdgeMap:
    """directed edge (u,v) -> (triangle index, third vertex): get_adjacent of DelaunayTriangulation"""

    def __init__(self, tri):
        T = tri.triangles.astype(np.int64)
        N = tri.num_points
        e = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
        self.N = N
        keys = e[:, 0] * N + e[:, 1]
        self.order = np.argsort(keys)
        self.keys = keys[self.order]
        self.tri = np.tile(np.arange(len(T)), 3)[self.order]
        self.third = np.concatenate([T[:, 2], T[:, 0], T[:, 1]])[self.order]

    def lookup(self, u, v):
        k = np.asarray(u, np.int64) * self.N + np.asarray(v, np.int64)
        pos = np.minimum(np.searchsorted(self.keys, k), len(self.keys) - 1)
        hit = self.keys[pos] == k
        return hit, np.where(hit, self.tri[pos], -1), np.where(hit, self.third[pos], -1)


def pl_interpolate(p, T, u, x, y):
    """pl_interpolate(prob, T, u, x, y) (utils.jl:23-27): piecewise-linear interpolant of u at points
    (x, y) inside the triangles with caller indices T (arrays broadcast); (n,) or (n, neq)."""
    T = np.atleast_1d(np.asarray(T, np.int32))
    xy = L.f64(np.stack(np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float)), -1).reshape(-1, 2))
    T = L.i32(np.broadcast_to(T, (len(xy),)))
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty((len(xy), p.engine.neq))
    L.check(p.engine.h, L.lib().fvm_eval_points(p.engine.h, 0.0, u.ctypes.data, 0, len(xy), L.ip(T), L.dp(xy), None, L.dp(out)))
    return out[:, 0] if p.prob.neqs == 0 else out


def compute_flux(p, i, j, u, t):
    """compute_flux(prob, i, j, u, t) (problem.jl:458-487): q(midpoint) . n for the edges (i, j), n the
    clockwise rotation of the edge (pointing right of i -> j), evaluated with the shape function of
    the triangle on that side (or the other side for boundary edges).  Arrays of edges are accepted."""
    mesh = p.prob.mesh
    tri = mesh.triangulation
    if getattr(mesh, "_edge_map", None) is None:
        mesh._edge_map = _EdgeMap(tri)
    i = np.atleast_1d(np.asarray(i, np.int64))
    j = np.atleast_1d(np.asarray(j, np.int64))
    P = tri.points
    e = P[j] - P[i]
    ell = np.sqrt(e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1])
    nrm = L.f64(np.stack([e[:, 1] / ell, -e[:, 0] / ell], 1))
    hit, t_right, _ = mesh._edge_map.lookup(j, i)      # the vertex in the direction of the normal
    hit2, t_left, _ = mesh._edge_map.lookup(i, j)
    if not (hit | hit2).all():
        raise KeyError("compute_flux: (i, j) is not an edge of the triangulation")
    tq = L.i32(np.where(hit, t_right, t_left))
    mid = L.f64((P[i] + P[j]) / 2)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty((len(i), p.engine.neq))
    L.check(p.engine.h, L.lib().fvm_eval_points(p.engine.h, float(t), u.ctypes.data, 0, len(i), L.ip(tq), L.dp(mid), L.dp(nrm), L.dp(out)))
    return out[:, 0] if p.prob.neqs == 0 else out
