"""The linear templates of /root/reference/src/specific_problems/*.jl on the GPU.

Constructors keep the reference's names and keyword arguments.  The reference builds a dense
`zeros(n, n)` and calls `sparse`; here A (CSR) and b are assembled once on the device
(fvm_assemble) and `solve` runs the device-resident fixed-step Tsit5 (transient templates) or the
Jacobi-preconditioned Krylov solvers (steady templates).  Template coefficient functions are
(x,y)-only, as in the reference (`diffusion_function(x, y, p)`, BC functions are called with
`t = u = nothing`, abstract_templates.jl:112-114), so plain NumPy-broadcastable callables are
accepted: they are tabulated on the host at setup."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import functors as F
from .conditions import Conditions, InternalConditions
from .problem import Engine

TPL_DIFFUSION, TPL_LINEAR_REACTION_DIFFUSION, TPL_MEAN_EXIT_TIME, TPL_POISSON, TPL_LAPLACE = range(5)
KRYLOV_PCG, KRYLOV_BICGSTAB = 0, 1


def _eval_cond(fn, x, y):
    """a(x, y, nothing, nothing, p) on arrays"""
    x = np.asarray(x, dtype=np.float64)
    if isinstance(fn, (int, float)):
        return np.full(x.shape, float(fn))
    if isinstance(fn, F.Const):
        return np.full(x.shape, fn.c)
    if isinstance(fn, F.LinearXY):
        return fn.c0 + fn.cx * x + fn.cy * np.asarray(y)
    if isinstance(fn, F.ExpXYT) and fn.ct == 0.0:
        return fn.c0 * np.exp(fn.cx * x + fn.cy * np.asarray(y))
    if isinstance(fn, (F.AffineU, F.ExpSaturation, F.ExpXYT)):
        raise TypeError("template conditions are evaluated with t = u = nothing; %r depends on t or u" % (fn,))
    return np.broadcast_to(np.asarray(fn(x, y, None, None, None), dtype=np.float64), x.shape)


def _eval_xy(fn, x, y, p):
    if isinstance(fn, (int, float)):
        return np.full(np.shape(x), float(fn))
    return np.broadcast_to(np.asarray(fn(x, y, p), dtype=np.float64), np.shape(x))


class Tsit5:
    """Device-resident Tsit5.  `Tsit5(dt)` is `solve(prob, Tsit5(); adaptive=false, dt=dt)`, the
    fixed-step integrator of north_star (c); `Tsit5()` / `Tsit5(abstol=, reltol=)` is the adaptive
    `solve(prob, Tsit5(); abstol, reltol, saveat)` with OrdinaryDiffEq's default tolerances."""

    def __init__(self, dt=None, adaptive=None, abstol=1e-6, reltol=1e-3):
        self.adaptive = (dt is None) if adaptive is None else bool(adaptive)
        if not self.adaptive and dt is None:
            raise ValueError("fixed-step Tsit5 needs dt")
        self.dt = None if dt is None else float(dt)
        self.abstol, self.reltol = float(abstol), float(reltol)
        self.naccept = self.nreject = None


def run_tsit5(h, alg, use_operator, u, t0, t1, ts, us):
    """Dispatches to the fixed-step or the adaptive device stepper; u is updated in place."""
    lib = L.lib()
    nts = len(ts)
    if alg.adaptive:
        na, nr = C.c_int64(), C.c_int64()
        L.check(h, lib.fvm_tsit5_adaptive(h, 1 if use_operator else 0, u.ctypes.data, t0, t1, alg.abstol, alg.reltol,
                                          alg.dt or 0.0, nts, L.dp(ts) if nts else None, us.ctypes.data if nts else None, 0,
                                          C.byref(na), C.byref(nr)))
        alg.naccept, alg.nreject = na.value, nr.value
    else:
        L.check(h, lib.fvm_tsit5(h, 1 if use_operator else 0, u.ctypes.data, t0, t1, alg.dt, nts, L.dp(ts) if nts else None,
                                 us.ctypes.data if nts else None, 0))


class KrylovJacobi:
    """Jacobi-preconditioned CG ("pcg") / BiCGStab ("bicgstab"); stands in for KLUFactorization."""

    def __init__(self, method=None, rtol=1e-12, maxiter=100000):
        self.method, self.rtol, self.maxiter = method, float(rtol), int(maxiter)


class Solution:
    def __init__(self, u, t=None, iters=None, relres=None, retcode="Success"):
        self.u, self.t, self.iters, self.relres, self.retcode = u, t, iters, relres, retcode


class AbstractFVMTemplate:
    template_id = None
    steady = False

    def _setup(self, mesh, BCs, ICs, diffusion_function, diffusion_parameters, source_function=None, source_parameters=None,
               tile_triangles=0, reference_quirks=True, ghost=None):
        self.mesh = mesh
        self.conditions = Conditions(mesh, BCs, ICs or InternalConditions())
        self.diffusion_function, self.diffusion_parameters = diffusion_function, diffusion_parameters
        self.source_function, self.source_parameters = source_function, source_parameters
        c = self.conditions
        err = None
        if self.steady and c.has_dudt_nodes():  # poissons_equation.jl:69-70, mean_exit_time.jl:66-67
            err = "%s does not support Dudt nodes." % ("MeanExitTimeProblem" if self.template_id == TPL_MEAN_EXIT_TIME else "PoissonsEquation")
        elif self.template_id == TPL_MEAN_EXIT_TIME and c.has_constrained_edges():  # mean_exit_time.jl:68-69
            err = "MeanExitTimeProblem does not support Constrained edges."
        if err and ghost is None:
            raise ValueError(err)
        # sharded: these checks see the rank-LOCAL conditions only.  A rank that raised here while its peers went on
        # into the NCCL calls of install_halo would hang them, so the verdict is deferred: install_halo combines it over
        # all ranks and every rank raises the same ValueError
        self._validation_error = err
        tri = mesh.triangulation
        P = tri.points
        N = tri.num_points
        uv = c.boundary_edges
        # tabulate the (x,y)-only coefficient functions
        d_const, d_edge, d_bnd = 1.0, None, None
        if isinstance(diffusion_function, (int, float)):
            d_const = float(diffusion_function)
        else:
            mid = mesh.cv_edge_midpoints()
            d_edge = L.f64(_eval_xy(diffusion_function, mid[..., 0], mid[..., 1], diffusion_parameters))
            qp = mesh.boundary_quarter_points(uv) if len(uv) else np.zeros((0, 2, 2))
            d_bnd = L.f64(_eval_xy(diffusion_function, qp[..., 0], qp[..., 1], diffusion_parameters))
        node_value = np.zeros(N)
        for fidx in np.unique(c.node_fidx[c.node_kind != 0]):
            sel = (c.node_kind != 0) & (c.node_fidx == fidx)
            if self.template_id == TPL_MEAN_EXIT_TIME:
                continue  # BC functions are never evaluated (mean_exit_time.jl:72-74)
            node_value[sel] = _eval_cond(c.functions[fidx], P[sel, 0], P[sel, 1])
        edge_value = np.zeros((len(uv), 2))
        if len(uv) and self.template_id != TPL_MEAN_EXIT_TIME:
            qp = mesh.boundary_quarter_points(uv)
            for fidx in np.unique(c.edge_fidx[c.edge_kind == 1]):
                sel = (c.edge_kind == 1) & (c.edge_fidx == fidx)
                edge_value[sel] = _eval_cond(c.functions[fidx], qp[sel, :, 0], qp[sel, :, 1])
        source = None
        if source_function is not None:
            source = L.f64(_eval_xy(source_function, P[:, 0], P[:, 1], source_parameters))
        # big tiles: the sliced-ELL SpMV keeps only x of a tile in shared memory, and fewer interface rows
        # (3 % at 4096 triangles per tile) mean less gather traffic in the tail kernel
        self.engine = Engine(mesh, 1, [c], tile_triangles=tile_triangles or 4096, geometry_mode=1, ghost=ghost)  # assembly recomputes geometry; no SoA kept
        self.engine.validation_error = err
        self.node_value = node_value
        self.N = N
        if err:
            return
        node_value, edge_value = L.f64(node_value), L.f64(edge_value)
        L.check(self.engine.h, L.lib().fvm_assemble(self.engine.h, self.template_id, d_const, L.dp(d_edge), L.dp(d_bnd),
                                                    L.dp(node_value), L.dp(edge_value), L.dp(source), 1 if reference_quirks else 0))
        self.N = N

    # A and b in the caller's numbering (scipy CSR), for parity checks and host-side use
    @property
    def A(self):
        return self._csr()[0]

    @property
    def b(self):
        return self._csr()[1]

    def _csr(self):
        if getattr(self, "_csr_cache", None) is None:
            import scipy.sparse as sp
            n, nnz = C.c_int64(), C.c_int64()
            L.check(self.engine.h, L.lib().fvm_get_csr_size(self.engine.h, C.byref(n), C.byref(nnz)))
            rowptr = np.empty(n.value + 1, np.int32)
            col = np.empty(nnz.value, np.int32)
            val = np.empty(nnz.value)
            b = np.empty(n.value)
            L.check(self.engine.h, L.lib().fvm_get_csr(self.engine.h, L.ip(rowptr), L.ip(col), L.dp(val), L.dp(b)))
            self._csr_cache = (sp.csr_matrix((val, col, rowptr), shape=(n.value, n.value)), b)
        return self._csr_cache

    def mul(self, du, u, add_b=True):
        """mul!(du, Aop, u) of the MatrixOperator (diffusion_equation.jl:93-94): du = A u + b"""
        L.check(self.engine.h, L.lib().fvm_spmv(self.engine.h, u.ctypes.data, du.ctypes.data, 1 if add_b else 0, 0))
        return du


class DiffusionEquation(AbstractFVMTemplate):
    """diffusion_equation.jl:69-101"""
    template_id = TPL_DIFFUSION

    def __init__(self, mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None, initial_condition,
                 initial_time=0.0, final_time, **kw):
        self._setup(mesh, BCs, ICs, diffusion_function, diffusion_parameters, **kw)
        self.initial_condition = np.ascontiguousarray(initial_condition, dtype=np.float64)
        self.initial_time, self.final_time = float(initial_time), float(final_time)
        ic = self.initial_condition.copy()
        d = self.conditions.node_kind == 1
        ic[d] = self.node_value[d]  # apply_dirichlet_conditions!, abstract_templates.jl:109-117
        self.u0 = ic

    def __repr__(self):
        return "DiffusionEquation with %d nodes and time span (%s, %s)" % (
            int(self.mesh.triangulation.solid_vertex_mask().sum()), self.initial_time, self.final_time)


class LinearReactionDiffusionEquation(DiffusionEquation):
    """linear_reaction_diffusion_equations.jl:76-125"""
    template_id = TPL_LINEAR_REACTION_DIFFUSION

    def __init__(self, mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None, source_function,
                 source_parameters=None, initial_condition, initial_time=0.0, final_time, **kw):
        self._setup(mesh, BCs, ICs, diffusion_function, diffusion_parameters, source_function, source_parameters, **kw)
        self.initial_condition = np.ascontiguousarray(initial_condition, dtype=np.float64)
        self.initial_time, self.final_time = float(initial_time), float(final_time)
        ic = self.initial_condition.copy()
        d = self.conditions.node_kind == 1
        ic[d] = self.node_value[d]
        self.u0 = ic


class MeanExitTimeProblem(AbstractFVMTemplate):
    """mean_exit_time.jl:57-94"""
    template_id = TPL_MEAN_EXIT_TIME
    steady = True

    def __init__(self, mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None, **kw):
        self._setup(mesh, BCs, ICs, diffusion_function, diffusion_parameters, **kw)


class PoissonsEquation(AbstractFVMTemplate):
    """poissons_equation.jl:58-88"""
    template_id = TPL_POISSON
    steady = True

    def __init__(self, mesh, BCs, ICs=None, *, diffusion_function=1.0, diffusion_parameters=None, source_function,
                 source_parameters=None, **kw):
        self._setup(mesh, BCs, ICs, diffusion_function, diffusion_parameters, source_function, source_parameters, **kw)


class LaplacesEquation(AbstractFVMTemplate):
    """laplaces_equation.jl:51-78"""
    template_id = TPL_LAPLACE
    steady = True

    def __init__(self, mesh, BCs, ICs=None, *, diffusion_function=1.0, diffusion_parameters=None, **kw):
        self._setup(mesh, BCs, ICs, diffusion_function, diffusion_parameters, **kw)


def solve_template(prob, alg=None, saveat=None, x0=None):
    """solve(prob::AbstractFVMTemplate, alg) (abstract_templates.jl:58-60)."""
    lib, h = L.lib(), prob.engine.h
    if prob.steady:
        alg = alg or KrylovJacobi()
        method = alg.method
        if method is None:  # PCG needs a symmetric operator: constant D and no Constrained edges
            sym = isinstance(prob.diffusion_function, (int, float)) and not prob.conditions.has_constrained_edges()
            method = "pcg" if sym else "bicgstab"
        x = np.zeros(prob.N) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        it, rr = C.c_int32(), C.c_double()
        L.check(h, lib.fvm_krylov(h, KRYLOV_PCG if method == "pcg" else KRYLOV_BICGSTAB, x.ctypes.data, alg.rtol, alg.maxiter,
                                  C.byref(it), C.byref(rr), 0))
        return Solution(x, None, it.value, rr.value, "Success" if rr.value <= 10 * alg.rtol else "MaxIters")
    if not isinstance(alg, Tsit5):
        raise TypeError("transient templates are integrated with the device-resident fixed-step Tsit5(dt)")
    u = prob.u0.copy()
    ts = np.ascontiguousarray([] if saveat is None else saveat, dtype=np.float64)
    us = np.full((len(ts), prob.N), np.nan)  # a row the stepper never wrote must not look like data
    run_tsit5(h, alg, True, u, prob.initial_time, prob.final_time, ts, us)
    # the reference's state is augmented by a trailing 1 that carries b (diffusion_equation.jl:82-94)
    if saveat is None:
        return Solution(np.append(u, 1.0), prob.final_time)
    return Solution([np.append(r, 1.0) for r in us], ts)
