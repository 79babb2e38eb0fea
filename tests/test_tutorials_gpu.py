"""Two of the reference's tutorials with closed-form solutions, end to end on the device (the tutorials themselves
only compare images): the Helmholtz problem solved with NewtonRaphson on a SteadyFVMProblem, and the porous-medium
equation against the Barenblatt profile.  The CPU oracle solves the same problems in tests/test_oracle_golden.py."""
import math

import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import RTOL_TSIT5, Pair, rel_err

pytestmark = pytest.mark.gpu


def test_helmholtz_steady_newton_raphson():
    """docs/src/literate_tutorials/helmholtz_equation_with_inhomogeneous_boundary_conditions.jl:20-65,82-84:
    solve(SteadyFVMProblem(prob), NewtonRaphson()) with q.n = -1 and source u; exact
    u = -(cos(x+1)+cos(1-x)+cos(y+1)+cos(1-y))/sin(2).  At the tutorial's 125x125 mesh against the closed form, at
    25x25 against the oracle's own steady state (one exact Newton step on its column-by-column Jacobian)."""
    def build(n):
        pair = Pair(G.triangulate_rectangle(-1, 1, -1, 1, n, n, single_boundary=True))
        gp, op = pair.problem(G.Const(-1.0), G.Neumann, G.ConstantDiffusion(1.0), source=G.LinearSource(1.0, 0.0),
                              ic=np.zeros(n * n), final_time=np.inf)
        return pair, gp, op

    pair, gp, op = build(125)
    sol = G.solve(G.SteadyFVMProblem(gp), G.NewtonRaphson())
    assert sol.retcode == "Success" and sol.iters <= 2  # affine problem: one Newton step
    x, y = pair.gtri.points[:, 0], pair.gtri.points[:, 1]
    exact = -(np.cos(x + 1) + np.cos(1 - x) + np.cos(y + 1) + np.cos(1 - y)) / math.sin(2)
    assert rel_err(sol.u, exact) <= 1e-4  # O(h^2): 1.0e-3 at 25x25 -> ~4e-5 at 125x125
    pair, gp, op = build(25)
    N = 625
    f0 = O.fvm_eqs_vec(np.zeros(N), np.zeros(N), op, 0.0).copy()
    J = np.stack([O.fvm_eqs_vec(np.zeros(N), e, op, 0.0) - f0 for e in np.eye(N)], axis=1)
    uref = np.linalg.solve(J, -f0)
    sol = G.solve(G.SteadyFVMProblem(gp), G.NewtonRaphson(), tile_triangles=128)
    assert rel_err(sol.u, uref) <= 1e-10
    # the device Jacobian is the oracle's Jacobian
    Jd = G.jacobian(np.zeros(N), G.get_cuda_parameters(gp, tile_triangles=128), 0.0).toarray()
    assert np.abs(Jd - J).max() <= 1e-11 * np.abs(J).max()


def test_porous_medium_barenblatt():
    """docs/src/literate_tutorials/porous_medium_equation.jl:20-50,88-99 (m = 2, M = 0.37, D = 2.53), started from
    the exact profile at t = 1: device Tsit5 with the Dirichlet callback vs the oracle's integrator (1e-10) and vs
    the Barenblatt solution at t = 2 (discretisation level)."""
    m, M, D = 2, 0.37, 2.53
    RmM = 4 * m / (m - 1) * (M / (4 * np.pi)) ** ((m - 1) / m)

    def exact(x, y, t):
        r2 = x * x + y * y
        inner = (M / (4 * np.pi)) ** ((m - 1) / m) - (m - 1) / (4 * m) * r2 * (D * t) ** (-1 / m)
        return np.where(r2 < RmM * (D * t) ** (1 / m), (D * t) ** (-1 / m) * np.maximum(inner, 0.0) ** (1 / (m - 1)), 0.0)

    pair = Pair(G.triangulate_rectangle(-3.0, 3.0, -3.0, 3.0, 61, 61, single_boundary=True))
    x, y = pair.gtri.points[:, 0], pair.gtri.points[:, 1]
    ic = exact(x, y, 1.0)
    gp = G.FVMProblem(pair.gmesh, G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet), diffusion_function=G.PowerDiffusion(D, float(m)),
                      initial_condition=ic, initial_time=1.0, final_time=2.0)
    op = O.FVMProblem(pair.omesh, O.BoundaryConditions(pair.omesh, lambda x, y, t, u, p: 0.0 * u, O.Dirichlet),
                      diffusion_function=lambda x, y, t, u, p: p[0] * u ** (p[1] - 1), diffusion_parameters=(D, m),
                      initial_condition=ic, initial_time=1.0, final_time=2.0)
    sol = G.solve(gp, G.Tsit5(0.005), tile_triangles=256)
    uref = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, op, t), ic, 1.0, 2.0, 0.005,
                         callback=lambda v, t: (O.update_dirichlet_nodes(v, t, op), True)[1])
    assert rel_err(sol.u, uref) <= RTOL_TSIT5
    ref = exact(x, y, 2.0)
    assert np.abs(sol.u - ref).max() <= 0.03 * ref.max()
    assert abs(sol.u @ pair.omesh.cv_volumes - M) <= 2e-3 * M


def test_wedge_mixed_conditions_bessel_series():
    """docs/src/literate_tutorials/diffusion_equation_in_a_wedge_with_mixed_boundary_conditions.jl:20-45: zero-flux Neumann
    edges, Dirichlet arc, f = 1 - r.  The device Tsit5 on fvm_eqs! with the Dirichlet callback against the oracle's
    integrator (1e-10), and both the FVMProblem and the DiffusionEquation template (device operator Tsit5) against the
    tutorial's exact Bessel series (discretisation level; the oracle itself is pinned to it on the CPU)."""
    from tests.common import wedge_mesh
    from tests.test_oracle_golden import _wedge_series
    pair = Pair(wedge_mesh(24))
    r = np.hypot(pair.gtri.points[:, 0], pair.gtri.points[:, 1])
    specs, types = (G.Const(0.0),) * 3, (G.Neumann, G.Dirichlet, G.Neumann)
    gp, op = pair.problem(specs, types, G.ConstantDiffusion(1.0), ic=1 - r, final_time=0.1)
    sol = G.solve(gp, G.Tsit5(2e-4), tile_triangles=128)
    uref = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, op, t), 1 - r, 0.0, 0.1, 2e-4,
                         callback=lambda v, t: (O.update_dirichlet_nodes(v, t, op), True)[1])
    assert rel_err(sol.u, uref) <= RTOL_TSIT5
    exact = _wedge_series(r, 0.1)
    assert np.abs(sol.u - exact).max() <= 4e-3 * exact.max()
    tpl = G.DiffusionEquation(pair.gmesh, G.BoundaryConditions(pair.gmesh, specs, types), diffusion_function=1.0,
                              initial_condition=1 - r, final_time=0.1, tile_triangles=128)
    tsol = G.solve(tpl, G.Tsit5(2e-4))
    assert len(tsol.u) == len(r) + 1 and tsol.u[-1] == 1.0  # the augmented state [u; 1] (diffusion_equation.jl:82-94)
    assert np.abs(tsol.u[:-1] - exact).max() <= 4e-3 * exact.max()
    assert rel_err(tsol.u[:-1], sol.u) <= 1e-9  # same semi-discrete system, integrated with and without the callback


def test_disk_dudt_boundary_exact_solution():
    """docs/src/literate_tutorials/reaction_diffusion_equation_with_a_time_dependent_dirichlet_boundary_condition_on_a_disk.jl
    :20-45: u_t = div(u grad u) + u(1-u), du/dt = u on the boundary, exact u = exp(t) sqrt(I0(sqrt(2) r)): Dudt nodes,
    u-dependent diffusion and a nonlinear source on an unstructured mesh, device Tsit5 vs the oracle's integrator and vs
    the exact solution."""
    from scipy.special import i0
    from tests.common import disk_mesh
    pair = Pair(disk_mesh(16))
    r = np.hypot(pair.gtri.points[:, 0], pair.gtri.points[:, 1])
    ic = np.sqrt(i0(np.sqrt(2) * r))
    gp, op = pair.problem(G.AffineU(0.0, 1.0), G.Dudt, G.PowerDiffusion(1.0, 2.0), source=G.LogisticSource(1.0), ic=ic, final_time=0.1)
    du = G.fvm_eqs(np.zeros_like(ic), ic, G.get_cuda_parameters(gp, tile_triangles=128), 0.0)
    assert rel_err(du, O.fvm_eqs_vec(np.zeros_like(ic), ic, op, 0.0)) <= 1e-12
    sol = G.solve(gp, G.Tsit5(5e-4), tile_triangles=128)
    uref = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, op, t), ic, 0.0, 0.1, 5e-4)
    assert rel_err(sol.u, uref) <= RTOL_TSIT5
    assert np.abs(sol.u - math.exp(0.1) * ic).max() <= 2e-4 * math.exp(0.1) * ic.max()
