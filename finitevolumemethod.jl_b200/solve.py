"""solve(...) glue (/root/reference/src/solve.jl:167-220, abstract_templates.jl:58-60)."""
import numpy as np

from . import _lib as L
from .problem import FVMProblem, FVMSystem, SteadyFVMProblem, get_cuda_parameters
from .templates import AbstractFVMTemplate, Solution, Tsit5, run_tsit5, solve_template


def solve(prob, alg=None, *, saveat=None, parallel="cuda", p=None, **kw):
    """`solve(prob, alg; saveat, parallel)`.

    * templates: device Tsit5 (transient) or Jacobi-Krylov (steady);
    * FVMProblem / FVMSystem with `Tsit5(dt)`: device-resident fixed-step Tsit5 on `fvm_eqs!` with the
      Dirichlet callback after every step (solve.jl:133-165);
    * any other integrator stays on the host and calls `fvm_eqs(du, u, p, t)` through
      `get_cuda_parameters(prob)`, exactly like the reference's ODEProblem(prob).f."""
    if isinstance(prob, AbstractFVMTemplate):
        return solve_template(prob, alg, saveat=saveat, **kw)
    if isinstance(prob, SteadyFVMProblem):
        raise NotImplementedError("SteadyFVMProblem is solved by a host nonlinear solver around fvm_eqs(du,u,p,t) "
                                  "(solve.jl:209-220); use get_cuda_parameters(prob) and your solver of choice")
    if not isinstance(prob, (FVMProblem, FVMSystem)):
        raise TypeError("cannot solve %r" % (prob,))
    if parallel != "cuda":
        raise ValueError("this package only provides the CUDA path (no CPU fallback)")
    if not isinstance(alg, Tsit5):
        raise TypeError("only Tsit5 runs on the device; other integrators call fvm_eqs through get_cuda_parameters")
    p = p or get_cuda_parameters(prob, **kw)
    u = np.ascontiguousarray(prob.initial_condition, dtype=np.float64).copy()
    ts = np.ascontiguousarray([] if saveat is None else saveat, dtype=np.float64)
    us = np.empty((len(ts),) + u.shape)
    run_tsit5(p.engine.h, alg, False, u, prob.initial_time, prob.final_time, ts, us)
    if saveat is None:
        return Solution(u, prob.final_time)
    return Solution(list(us), ts)
