"""solve(...) glue (/root/reference/src/solve.jl:167-220, abstract_templates.jl:58-60)."""
import numpy as np

from . import _lib as L
from .problem import FVMProblem, FVMSystem, SteadyFVMProblem, fvm_eqs, get_cuda_parameters, jacobian, update_dirichlet_nodes
from .templates import AbstractFVMTemplate, Solution, Tsit5, run_tsit5, solve_template


class NewtonRaphson:
    """`solve(SteadyFVMProblem(prob), NewtonRaphson())` (solve.jl:209-220; used by
    docs/src/literate_tutorials/helmholtz_equation_with_inhomogeneous_boundary_conditions.jl:65).

    linsolve="bicgstab" (default): the whole iteration runs on the device (`fvm_newton`): fvm_eqs!, its sparse
    Jacobian and a Jacobi-preconditioned BiCGStab on the Jacobian's block CSR; only the residual norm crosses PCIe.
    linsolve="direct": the Jacobian is fetched every iteration and factorised on the host (SciPy SuperLU, like the
    reference's NonlinearSolve + KLU) -- for systems the Jacobi-Krylov solve does not converge on."""

    def __init__(self, abstol=1e-11, reltol=1e-11, maxiters=50, linsolve="bicgstab", lin_rtol=1e-13, lin_maxiters=20000):
        if linsolve not in ("bicgstab", "direct"):
            raise ValueError("NewtonRaphson: linsolve must be 'bicgstab' (device) or 'direct' (host SuperLU)")
        self.abstol, self.reltol, self.maxiters = float(abstol), float(reltol), int(maxiters)
        self.linsolve, self.lin_rtol, self.lin_maxiters = linsolve, float(lin_rtol), int(lin_maxiters)


def _solve_steady_newton_device(prob, alg, p):
    import ctypes as C
    u = np.ascontiguousarray(prob.initial_condition, dtype=np.float64).copy()
    it, lin = C.c_int32(), C.c_int64()
    res, r0 = C.c_double(), C.c_double()
    h = p.engine.h
    L.check(h, L.lib().fvm_newton(h, prob.initial_time, u.ctypes.data, alg.abstol, alg.reltol, alg.maxiters, alg.lin_rtol, alg.lin_maxiters,
                                  C.byref(it), C.byref(res), C.byref(r0), C.byref(lin), 0))
    ok = res.value <= alg.abstol + alg.reltol * r0.value
    sol = Solution(u, None, iters=it.value, relres=(res.value / r0.value) if r0.value > 0 else 0.0, retcode="Success" if ok else "MaxIters")
    sol.linear_iters = lin.value
    return sol


def _solve_steady_newton(prob, alg, p):
    if alg.linsolve == "bicgstab":
        return _solve_steady_newton_device(prob, alg, p)
    import scipy.sparse.linalg as spla
    t = prob.initial_time
    u = np.ascontiguousarray(prob.initial_condition, dtype=np.float64).copy()
    update_dirichlet_nodes(u, t, p)  # Dirichlet rows of fvm_eqs! are identically zero: fix the values first
    du = np.zeros_like(u)
    fvm_eqs(du, u, p, t)
    r0 = np.abs(du).max()
    retcode, it = "MaxIters", 0
    for it in range(alg.maxiters + 1):
        res = np.abs(du).max()
        if res <= alg.abstol + alg.reltol * r0:
            retcode = "Success"
            break
        if it == alg.maxiters or not np.isfinite(res):
            break
        J = jacobian(u, p, t).tocsr()
        # rows without any entry (Dirichlet nodes, points that are not vertices) have du = 0 identically:
        # those unknowns keep their value and the Newton system is solved on the remaining block
        free = np.flatnonzero(np.asarray(abs(J).sum(axis=1)).ravel() != 0.0)
        delta = np.zeros(J.shape[0])
        delta[free] = spla.splu(J[free][:, free].tocsc()).solve(-du.ravel()[free])
        u += delta.reshape(u.shape)
        fvm_eqs(du, u, p, t)
    return Solution(u, None, iters=it, relres=float(np.abs(du).max() / r0) if r0 > 0 else 0.0, retcode=retcode)


def solve(prob, alg=None, *, saveat=None, parallel="cuda", p=None, **kw):
    """`solve(prob, alg; saveat, parallel)`.

    * templates: device Tsit5 (transient) or Jacobi-Krylov (steady);
    * SteadyFVMProblem with `NewtonRaphson()`: host Newton iteration on the device RHS and device Jacobian;
    * FVMProblem / FVMSystem with `Tsit5(dt)`: device-resident fixed-step Tsit5 on `fvm_eqs!` with the
      Dirichlet callback after every step (solve.jl:133-165);
    * any other integrator stays on the host and calls `fvm_eqs(du, u, p, t)` through
      `get_cuda_parameters(prob)`, exactly like the reference's ODEProblem(prob).f."""
    if isinstance(prob, AbstractFVMTemplate):
        return solve_template(prob, alg, saveat=saveat, **kw)
    if isinstance(prob, SteadyFVMProblem):
        if not isinstance(alg, NewtonRaphson):
            raise TypeError("SteadyFVMProblem: NewtonRaphson() is provided; any other nonlinear / DynamicSS solver stays on "
                            "the host around fvm_eqs(du,u,p,t) and jacobian(u,p,t) (solve.jl:209-220)")
        if parallel != "cuda":
            raise ValueError("this package only provides the CUDA path (no CPU fallback)")
        return _solve_steady_newton(prob.problem, alg, p or get_cuda_parameters(prob, **kw))
    if not isinstance(prob, (FVMProblem, FVMSystem)):
        raise TypeError("cannot solve %r" % (prob,))
    if parallel != "cuda":
        raise ValueError("this package only provides the CUDA path (no CPU fallback)")
    if not isinstance(alg, Tsit5):
        raise TypeError("only Tsit5 runs on the device; other integrators call fvm_eqs through get_cuda_parameters")
    p = p or get_cuda_parameters(prob, **kw)
    u = np.ascontiguousarray(prob.initial_condition, dtype=np.float64).copy()
    ts = np.ascontiguousarray([] if saveat is None else saveat, dtype=np.float64)
    us = np.full((len(ts),) + u.shape, np.nan)  # a row the stepper never wrote must not look like data
    run_tsit5(p.engine.h, alg, False, u, prob.initial_time, prob.final_time, ts, us)
    if saveat is None:
        return Solution(u, prob.final_time)
    return Solution(list(us), ts)
