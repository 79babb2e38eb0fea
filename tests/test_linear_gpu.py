"""GPU parity of the linear-template path: device assembly (CSR A, b), SpMV, device Tsit5 and the
Jacobi-Krylov steady solves, against the CPU oracle's restatement of
/root/reference/src/specific_problems/*.jl.  Tolerances: operator application 1e-12 (BASELINE.json),
fixed-step end state 1e-10, steady solves to the Krylov tolerance, closed forms at the reference's
own tolerances."""
import math

import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import RTOL_RHS, RTOL_TSIT5, Pair, cond_closure, delaunay_mesh, rel_err
from tests.test_rhs_gpu import _split_loop

pytestmark = pytest.mark.gpu


def xy_cond(spec):
    c = cond_closure(spec)
    return lambda x, y, t, u, p: c(x, y, 0.0, 0.0, p)


def direct_relres(ref):
    """relative residual the oracle's sparse-direct solve attains: the fp64 floor eps*|A||x|/|b|"""
    x = O.solve_steady(ref)
    return x, np.linalg.norm(ref.b - ref.A @ x) / np.linalg.norm(ref.b)


def assert_system_matches(tpl, ref, rtol=1e-12):
    A, b = tpl.A, tpl.b
    D = (A - ref.A).tocoo()
    scale = abs(ref.A).max()
    assert (abs(D.data).max() if D.nnz else 0.0) <= rtol * scale
    assert rel_err(b, ref.b) <= rtol if np.abs(ref.b).max() > 0 else np.abs(b).max() == 0


def test_diffusion_equation_readme_operator():
    """BASELINE configs[0]/[1] shape: DiffusionEquation on the README mesh; A, b, u0, A*u+b."""
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True))
    ic = np.where(pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    oBC = O.BoundaryConditions(pair.omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)
    tpl = G.DiffusionEquation(pair.gmesh, gBC, diffusion_function=1 / 9, initial_condition=ic, final_time=0.5)
    ref = O.DiffusionEquation(pair.omesh, oBC, diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.5)
    assert tpl.A.nnz == 17102  # structural pattern N + 2E
    assert_system_matches(tpl, ref)
    assert np.array_equal(tpl.u0, ref.u0)
    u = 50 * np.random.default_rng(20240517).random(len(ic))
    du = tpl.mul(np.empty_like(u), u)
    assert rel_err(du, ref.A @ u + ref.b) <= RTOL_RHS
    # the template operator equals the generic RHS (docs/src/literate_wyos/diffusion_equations.jl:436-438)
    gp, _ = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), ic=ic)
    rhs = G.fvm_eqs(np.zeros_like(u), u, G.get_cuda_parameters(gp), 0.0)
    assert rel_err(du, rhs) <= RTOL_RHS


@pytest.mark.parametrize("kind", ["diffusion", "lrd"])
def test_assembly_unstructured_mixed_conditions(kind):
    """Tabulated D(x,y), Neumann with a non-zero function, Dirichlet, Dudt and Constrained sections
    (with the reference's V_i quirk, abstract_templates.jl:262), internal conditions, points that are
    not vertices (fix_missing_vertices!)."""
    gtri = _split_loop(delaunay_mesh(1200, 21, extra_points=4))
    pair = Pair(gtri)
    specs = (G.LinearXY(0.5, 1.0, -2.0), G.Const(0.3), G.LinearXY(0.1, 0.2, 0.3), G.Const(0.0))
    types = (G.Dirichlet, G.Dudt, G.Neumann, G.Constrained)
    dn, tn = {200: 0, 201: 0}, {300: 0, 200: 0}
    gBC = G.BoundaryConditions(pair.gmesh, specs, types)
    oBC = O.BoundaryConditions(pair.omesh, tuple(xy_cond(s) for s in specs), types)
    gIC = G.InternalConditions((G.Const(0.7),), dirichlet_nodes=dn, dudt_nodes=tn)
    oIC = O.InternalConditions((xy_cond(G.Const(0.7)),), dirichlet_nodes=dn, dudt_nodes=tn)
    Dfn = lambda x, y, p: 1.0 + 0.5 * np.sin(3 * x) * np.cos(2 * y) + p
    Sfn = lambda x, y, p: -0.3 + x * y
    ic = np.random.default_rng(2).random(gtri.num_points)
    if kind == "diffusion":
        tpl = G.DiffusionEquation(pair.gmesh, gBC, gIC, diffusion_function=Dfn, diffusion_parameters=0.25,
                                  initial_condition=ic, final_time=1.0, tile_triangles=128)
        ref = O.DiffusionEquation(pair.omesh, oBC, oIC, diffusion_function=Dfn, diffusion_parameters=0.25,
                                  initial_condition=ic, final_time=1.0)
    else:
        tpl = G.LinearReactionDiffusionEquation(pair.gmesh, gBC, gIC, diffusion_function=Dfn, diffusion_parameters=0.25,
                                                source_function=Sfn, initial_condition=ic, final_time=1.0)
        ref = O.LinearReactionDiffusionEquation(pair.omesh, oBC, oIC, diffusion_function=Dfn, diffusion_parameters=0.25,
                                                source_function=Sfn, initial_condition=ic, final_time=1.0)
    assert_system_matches(tpl, ref)
    assert rel_err(tpl.u0, ref.u0) == 0.0
    u = np.random.default_rng(3).random(gtri.num_points)
    assert rel_err(tpl.mul(np.empty_like(u), u), ref.A @ u + ref.b) <= RTOL_RHS
    assert rel_err(tpl.mul(np.empty_like(u), u, add_b=False), ref.A @ u) <= RTOL_RHS


def test_tsit5_operator_readme():
    """DiffusionEquation + fixed-step Tsit5 to t = 0.5 (README, BASELINE configs[0]) vs the oracle's
    fixed-step Tsit5 on (A, b); end state within 1e-10, trailing augmented 1 (Appendix D-2)."""
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True))
    ic = np.where(pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    oBC = O.BoundaryConditions(pair.omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)
    tpl = G.DiffusionEquation(pair.gmesh, gBC, diffusion_function=1 / 9, initial_condition=ic, final_time=0.5)
    ref = O.DiffusionEquation(pair.omesh, oBC, diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.5)
    dt = 0.0025  # stable: dt < 3.3 h^2 / (8 D)
    sol = G.solve(tpl, G.Tsit5(dt), saveat=[0.0, 0.25, 0.5])

    def f(du, u, t):
        du[...] = ref.A @ u + ref.b

    uref, saves = O.tsit5_fixed(f, ref.u0, 0.0, 0.5, dt, saveat=[0.25, 0.5])
    assert len(sol.u) == 3 and all(len(s) == 2501 and s[-1] == 1.0 for s in sol.u)
    assert np.array_equal(sol.u[0][:-1], ref.u0)
    assert rel_err(sol.u[1][:-1], saves[0]) <= RTOL_TSIT5
    assert rel_err(sol.u[2][:-1], uref) <= RTOL_TSIT5
    end = G.solve(tpl, G.Tsit5(dt))
    assert np.array_equal(end.u, sol.u[2])
    assert 1.0 < uref.max() < 50.0  # the heat has diffused but not vanished


def test_tsit5_rhs_with_dirichlet_callback():
    """FVMProblem + device Tsit5: time-dependent Dirichlet values re-applied after every step
    (solve.jl:133-165), FSAL discarded after the callback; nonlinear diffusion + logistic source."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 24, 24, single_boundary=False))
    specs = (G.ExpSaturation(2.0, 0.5), G.Const(0.2), G.Const(0.0), G.AffineU(0.0, 1.0))
    types = (G.Dirichlet, G.Dirichlet, G.Neumann, G.Dudt)
    P = pair.gtri.points
    ic = 0.2 + 0.1 * np.sin(3 * P[:, 0]) * np.cos(2 * P[:, 1])
    gp, op = pair.problem(specs, types, G.PowerDiffusion(0.05, 2.0), source=G.LogisticSource(0.7), ic=ic, final_time=0.2)
    dt = 0.001
    sol = G.solve(gp, G.Tsit5(dt), saveat=[0.1, 0.2], tile_triangles=128)
    uref, saves = O.tsit5_fixed(lambda du, u, t: O.fvm_eqs_vec(du, u, op, t), ic, 0.0, 0.2, dt,
                                callback=lambda u, t: (O.update_dirichlet_nodes(u, t, op), True)[1], saveat=[0.1, 0.2])
    assert rel_err(sol.u[0], saves[0]) <= RTOL_TSIT5 and rel_err(sol.u[1], uref) <= RTOL_TSIT5
    assert abs(uref[0] - 2.0 * (1 - math.exp(-0.2 / 0.5))) < 1e-12  # the Dirichlet value at t = 0.2


def test_tsit5_system_no_callback_uses_fsal():
    """Gray-Scott FVMSystem, all-Neumann: no Dirichlet nodes, so k7 is reused as k1 (FSAL)."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 20, 17, single_boundary=True))
    N = pair.gtri.num_points
    rng = np.random.default_rng(4)
    U = np.ascontiguousarray(np.stack([0.5 + 0.5 * rng.random(N), 0.25 * rng.random(N)], axis=1))
    src = G.GrayScottSource(0.04, 0.1)
    g1, o1 = pair.problem(G.Const(0.0), G.Neumann, G.ConstantDiffusion(2e-3), source=src, var=0, ic=U[:, 0], final_time=2.0)
    g2, o2 = pair.problem(G.Const(0.0), G.Neumann, G.ConstantDiffusion(1e-3), source=src, var=1, ic=U[:, 1], final_time=2.0)
    gs, os_ = G.FVMSystem(g1, g2), O.FVMSystem(o1, o2)
    sol = G.solve(gs, G.Tsit5(0.05))
    uref = O.tsit5_fixed(lambda du, u, t: O.fvm_eqs_vec(du, u, os_, t), U, 0.0, 2.0, 0.05)
    assert sol.u.shape == (N, 2) and rel_err(sol.u, uref) <= RTOL_TSIT5


def test_poisson_closed_form_pcg():
    """docs/src/literate_wyos/poissons_equation.jl:92-116,159-162 on 100x100, plus parity with the
    oracle's sparse-direct solve (SuperLU standing in for KLU)."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 100, 100, single_boundary=True))
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    oBC = O.BoundaryConditions(pair.omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)
    src = lambda x, y, p: -np.sin(np.pi * x) * np.sin(np.pi * y)
    tpl = G.PoissonsEquation(pair.gmesh, gBC, source_function=src)
    ref = O.PoissonsEquation(pair.omesh, oBC, source_function=src)
    assert_system_matches(tpl, ref)
    sol = G.solve(tpl, G.KrylovJacobi("pcg", rtol=1e-13))
    P = pair.gtri.points
    exact = 1 / (2 * np.pi**2) * np.sin(np.pi * P[:, 0]) * np.sin(np.pi * P[:, 1])
    assert np.linalg.norm(sol.u - exact) <= 1e-4 * np.linalg.norm(exact)
    assert sol.relres <= 1e-12 and sol.iters < 2000
    assert rel_err(sol.u, O.solve_steady(ref)) <= 1e-9
    # BiCGStab reaches the same answer
    sol2 = G.solve(tpl, G.KrylovJacobi("bicgstab", rtol=1e-13))
    assert sol2.relres <= 1e-12 and rel_err(sol2.u, sol.u) <= 1e-8


def test_laplace_variable_diffusion_bicgstab():
    """docs/src/literate_wyos/laplaces_equation.jl:160-193: D = (x+1)(y+2), mixed Neumann/Dirichlet,
    exact solution 5 log6(1+x) at rtol 1e-3; non-symmetric operator -> BiCGStab."""
    pair = Pair(G.triangulate_rectangle(0, 5, 0, 5, 100, 100, single_boundary=False))
    specs = (G.Const(0.0), G.Const(5.0), G.Const(0.0), G.Const(0.0))
    types = (G.Neumann, G.Dirichlet, G.Neumann, G.Dirichlet)
    gBC = G.BoundaryConditions(pair.gmesh, specs, types)
    oBC = O.BoundaryConditions(pair.omesh, tuple(xy_cond(s) for s in specs), types)
    Dfn = lambda x, y, p: (x + 1) * (y + 2)
    tpl = G.LaplacesEquation(pair.gmesh, gBC, diffusion_function=Dfn)
    ref = O.LaplacesEquation(pair.omesh, oBC, diffusion_function=Dfn)
    assert_system_matches(tpl, ref)
    sol = G.solve(tpl, G.KrylovJacobi(rtol=1e-13))
    exact = 5 * np.log(1 + pair.gtri.points[:, 0]) / math.log(6)
    assert np.linalg.norm(sol.u - exact) <= 1e-3 * np.linalg.norm(exact)
    uref, floor = direct_relres(ref)
    assert sol.relres <= max(1e-12, 10 * floor)  # |b| is tiny here: 1e-12 |b| is below the fp64 floor
    assert rel_err(sol.u, uref) <= 1e-8


def test_mean_exit_time_unstructured():
    """mean_exit_time.jl:57-94: b = -1 on free vertices, identity rows on Dirichlet nodes, Neumann
    section ignored by the assembly, BC functions never evaluated."""
    gtri = _split_loop(delaunay_mesh(2500, 8, extra_points=2), k=2)
    pair = Pair(gtri)
    gBC = G.BoundaryConditions(pair.gmesh, (G.Const(99.0), G.Const(99.0)), (G.Neumann, G.Dirichlet))
    oBC = O.BoundaryConditions(pair.omesh, (xy_cond(G.Const(99.0)),) * 2, (O.Neumann, O.Dirichlet))
    tpl = G.MeanExitTimeProblem(pair.gmesh, gBC, diffusion_function=6.25e-4)
    ref = O.MeanExitTimeProblem(pair.omesh, oBC, diffusion_function=lambda x, y, p: 6.25e-4)
    assert_system_matches(tpl, ref)
    sol = G.solve(tpl, G.KrylovJacobi(rtol=1e-13))
    uref, floor = direct_relres(ref)
    assert sol.relres <= max(1e-12, 10 * floor)
    assert rel_err(sol.u, uref) <= 1e-9
    assert uref.max() > 10  # exit times are positive and large for a small D
    assert np.all(sol.u[-2:] == 0.0)  # points that are not vertices: A[i,i] = 1, b[i] = 0


def test_adaptive_tsit5_device():
    """solve(prob, Tsit5(); saveat) (README.md:43 style): adaptive device stepper vs the oracle's
    restatement of the same controller (same accepted/rejected counts, end state to round-off), and
    vs a fine fixed-step solution to the requested tolerance.  Template and FVMProblem paths."""
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True))
    ic = np.where(pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    oBC = O.BoundaryConditions(pair.omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)
    tpl = G.DiffusionEquation(pair.gmesh, gBC, diffusion_function=1 / 9, initial_condition=ic, final_time=0.5)
    ref = O.DiffusionEquation(pair.omesh, oBC, diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.5)
    alg = G.Tsit5(abstol=1e-8, reltol=1e-6)
    sol = G.solve(tpl, alg, saveat=[0.1, 0.5])

    def f(du, u, t):
        du[...] = ref.A @ u + ref.b

    uref, saves, nacc, nrej = O.tsit5_adaptive(f, ref.u0, 0.0, 0.5, 1e-8, 1e-6, saveat=[0.1, 0.5])
    # the step sequence is sensitive to the last bits of the error norm (summation order), so the counts
    # may differ by a step or two and the states agree to the tolerance level, not to round-off
    assert abs(alg.naccept - nacc) <= 2 and abs(alg.nreject - nrej) <= 2 and nacc > 20
    assert rel_err(sol.u[0][:-1], saves[0]) <= 1e-6 and rel_err(sol.u[1][:-1], uref) <= 1e-6
    fine = O.tsit5_fixed(f, ref.u0, 0.0, 0.5, 0.00125)
    assert rel_err(uref, fine) <= 1e-5 and rel_err(sol.u[1][:-1], fine) <= 1e-5
    # FVMProblem path with the Dirichlet callback after every accepted step
    gp, op = pair.problem(G.ExpSaturation(5.0, 0.2), G.Dirichlet, G.PowerDiffusion(0.02, 2.0), source=G.LogisticSource(0.5),
                          ic=1.0 + 0.01 * ic, final_time=0.3)
    alg2 = G.Tsit5(abstol=1e-7, reltol=1e-5)
    sol2 = G.solve(gp, alg2)
    u2, _, na2, nr2 = O.tsit5_adaptive(lambda d, x, t: O.fvm_eqs_vec(d, x, op, t), 1.0 + 0.01 * ic, 0.0, 0.3, 1e-7, 1e-5,
                                       callback=lambda x, t: (O.update_dirichlet_nodes(x, t, op), True)[1])
    assert abs(alg2.naccept - na2) <= 2 and abs(alg2.nreject - nr2) <= 2
    assert rel_err(sol2.u, u2) <= 1e-5


def test_annulus_two_boundary_loops():
    """A domain with a hole (docs/src/literate_tutorials/diffusion_equation_on_an_annulus.jl:54-68):
    inner loop Dirichlet 50(1-exp(-t/2)), outer loop zero-flux Neumann, D = 1.  RHS, device Tsit5 with
    the time-dependent Dirichlet callback, and Laplace's equation against ln(r/R)/ln(r0/R)."""
    from tests.common import annulus_mesh
    gtri = annulus_mesh(24, 96)
    assert len(gtri.boundary_sections) == 2
    inner_first = np.allclose(np.hypot(*gtri.points[gtri.boundary_sections[0][0]]), 0.2)
    pair = Pair(gtri)
    N = gtri.num_points
    specs = (G.ExpSaturation(50.0, 2.0), G.Const(0.0)) if inner_first else (G.Const(0.0), G.ExpSaturation(50.0, 2.0))
    types = (G.Dirichlet, G.Neumann) if inner_first else (G.Neumann, G.Dirichlet)
    u0 = 10.0 * np.exp(-25.0 * ((gtri.points[:, 0] + 0.5) ** 2 + (gtri.points[:, 1] + 0.5) ** 2))
    gp, op = pair.problem(specs, types, G.ConstantDiffusion(1.0), ic=u0, final_time=0.02)
    p = G.get_cuda_parameters(gp, tile_triangles=256)
    u = u0 + np.random.default_rng(1).random(N)
    assert rel_err(G.fvm_eqs(np.zeros(N), u, p, 0.3), O.fvm_eqs_vec(np.zeros(N), u, op, 0.3)) <= RTOL_RHS
    dt = 2e-5
    sol = G.solve(gp, G.Tsit5(dt), p=p)
    uref = O.tsit5_fixed(lambda d, x, t: O.fvm_eqs_vec(d, x, op, t), u0, 0.0, 0.02, dt,
                         callback=lambda x, t: (O.update_dirichlet_nodes(x, t, op), True)[1])
    assert rel_err(sol.u, uref) <= RTOL_TSIT5
    # steady: u = 1 on the inner circle, 0 on the outer one -> u(r) = ln(r/R)/ln(r0/R)
    one = (G.Const(1.0), G.Const(0.0)) if inner_first else (G.Const(0.0), G.Const(1.0))
    gBC = G.BoundaryConditions(pair.gmesh, one, (G.Dirichlet, G.Dirichlet))
    oBC = O.BoundaryConditions(pair.omesh, tuple(xy_cond(s) for s in one), (O.Dirichlet, O.Dirichlet))
    tpl = G.LaplacesEquation(pair.gmesh, gBC)
    ref = O.LaplacesEquation(pair.omesh, oBC)
    assert_system_matches(tpl, ref)
    s2 = G.solve(tpl, G.KrylovJacobi("pcg", rtol=1e-13))
    rr = np.hypot(gtri.points[:, 0], gtri.points[:, 1])
    exact = np.log(rr / 1.0) / np.log(0.2 / 1.0)
    assert rel_err(s2.u, O.solve_steady(ref)) <= 1e-9
    assert np.abs(s2.u - exact).max() <= 5e-3


def test_brusselator_tutorial_exact_solution():
    """docs/src/literate_tutorials/reaction_diffusion_brusselator_system_of_pdes.jl:95-140: a 2-species
    FVMSystem with per-species mixed, time-dependent Neumann / Dirichlet data and the exact solution
    Phi = exp(-x-y-t/2), Psi = exp(x+y+t/2).  Device Tsit5 vs the oracle (1e-10) and vs the exact
    solution (discretisation level)."""
    n = 31
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, n, n, single_boundary=False))
    P = pair.gtri.points
    E = G.ExpXYT
    phi_bc = (E(-0.25, -1, 0, -0.5), E(0.25 * np.exp(-1), 0, -1, -0.5), E(np.exp(-1), -1, 0, -0.5), E(-0.25, 0, -1, -0.5))
    psi_bc = (E(1.0, 1, 0, 0.5), E(-0.25 * np.exp(1), 0, 1, 0.5), E(-0.25 * np.exp(1), 1, 0, 0.5), E(1.0, 0, 1, 0.5))
    phi_t, psi_t = (G.Neumann, G.Neumann, G.Dirichlet, G.Neumann), (G.Dirichlet, G.Neumann, G.Neumann, G.Dirichlet)
    phi0, psi0 = np.exp(-P[:, 0] - P[:, 1]), np.exp(P[:, 0] + P[:, 1])
    src = G.BrusselatorSource()
    g1, o1 = pair.problem(phi_bc, phi_t, G.ConstantDiffusion(0.25), source=src, var=0, ic=phi0, final_time=0.5)
    g2, o2 = pair.problem(psi_bc, psi_t, G.ConstantDiffusion(0.25), source=src, var=1, ic=psi0, final_time=0.5)
    gs, os_ = G.FVMSystem(g1, g2), O.FVMSystem(o1, o2)
    U0 = gs.initial_condition
    du = G.fvm_eqs(np.zeros_like(U0), U0, G.get_cuda_parameters(gs, tile_triangles=256), 0.3)
    assert rel_err(du, O.fvm_eqs_vec(np.zeros_like(U0), U0, os_, 0.3)) <= RTOL_RHS
    dt = 1e-3
    sol = G.solve(gs, G.Tsit5(dt), tile_triangles=256)
    uref = O.tsit5_fixed(lambda d, x, t: O.fvm_eqs_vec(d, x, os_, t), U0, 0.0, 0.5, dt,
                         callback=lambda x, t: (O.update_dirichlet_nodes(x, t, os_), True)[1])
    assert rel_err(sol.u, uref) <= RTOL_TSIT5
    exact = np.stack([np.exp(-P[:, 0] - P[:, 1] - 0.25), np.exp(P[:, 0] + P[:, 1] + 0.25)], axis=1)
    assert np.abs(sol.u - exact).max() <= 2e-2 * np.abs(exact).max()


def test_steady_newton_raphson_linear_and_nonlinear():
    """solve(SteadyFVMProblem(prob), NewtonRaphson()) (solve.jl:209-220): device RHS + device Jacobian + device
    BiCGStab (fvm_newton); linsolve="direct" keeps the host sparse-direct solve.  Linear diffusion reaches the LaplacesEquation template solution (the reference's own cross-check,
    docs/src/literate_wyos/laplaces_equation.jl:188-193) in one step; the porous-medium problem is checked
    through the ORACLE's residual at the returned state."""
    gtri = _split_loop(delaunay_mesh(500, 13, jitter=0.3), k=4)
    pair = Pair(gtri)
    specs = (G.LinearXY(0.5, 1.0, -2.0), G.Const(0.0), G.LinearXY(0.1, 0.2, 0.3), G.Const(0.0))
    types = (G.Dirichlet, G.Neumann, G.Dirichlet, G.Neumann)
    ic = np.random.default_rng(1).random(gtri.num_points)
    gp, op = pair.problem(specs, types, G.ConstantDiffusion(0.8), ic=ic)
    sol = G.solve(G.SteadyFVMProblem(gp), G.NewtonRaphson())
    assert sol.retcode == "Success" and sol.iters <= 2
    oBC = O.BoundaryConditions(pair.omesh, tuple(xy_cond(s) for s in specs), types)
    ref = O.solve_steady(O.LaplacesEquation(pair.omesh, oBC, diffusion_function=lambda x, y, p: 0.8))
    assert rel_err(sol.u, ref) <= 1e-9
    # nonlinear: D = 0.3 u, source 0.5 - 0.2 u
    gp, op = pair.problem(specs, types, G.PowerDiffusion(0.3, 2.0), source=G.LinearSource(-0.2, 0.5), ic=0.5 + ic)
    sol = G.solve(G.SteadyFVMProblem(gp), G.NewtonRaphson(), tile_triangles=128)
    assert sol.retcode == "Success" and 2 <= sol.iters <= 20 and sol.linear_iters > sol.iters  # device Newton + BiCGStab (fvm_newton)
    direct = G.solve(G.SteadyFVMProblem(gp), G.NewtonRaphson(linsolve="direct"), tile_triangles=128)  # host SuperLU on the device Jacobian
    assert direct.retcode == "Success" and direct.iters == sol.iters and rel_err(sol.u, direct.u) <= 1e-9
    u0 = (0.5 + ic).copy()
    O.update_dirichlet_nodes(u0, 0.0, op)
    f0 = np.abs(O.fvm_eqs_vec(np.zeros_like(u0), u0, op, 0.0)).max()
    fs = np.abs(O.fvm_eqs_vec(np.zeros_like(sol.u), sol.u, op, 0.0)).max()
    assert fs <= 1e-9 * f0
    dn = np.array(sorted(op.conditions.dirichlet_nodes))
    assert rel_err(sol.u[dn], u0[dn]) <= 1e-15 and sol.u.min() > 0  # device FMA vs NumPy in c0 + cx x + cy y
    with pytest.raises(TypeError):
        G.solve(G.SteadyFVMProblem(gp), G.Tsit5(0.1))


@pytest.mark.parametrize("case", ["readme_50x50", "mean_exit_time_unstructured"])
def test_operator_host_buffer_pipeline_is_bit_identical(case):
    """fvm_spmv with host vectors takes the same banded pipeline as fvm_rhs (tiles own the interior rows, the
    sliced-ELL tail rows run slice by slice): forced on for small meshes, every band count must reproduce the
    plain schedule bit for bit, with and without b."""
    from tests.golden_cases import CASES, build_template
    from tests.test_rhs_gpu import _pipeline_env
    c = CASES[case](None)
    rng = np.random.default_rng(5)
    xs = [c.u, rng.random(len(c.u))]
    try:
        _pipeline_env(FVM_NO_PIPELINE=1)
        tpl = build_template(c, "gpu", tile_triangles=128)
        refs = [(tpl.mul(np.empty_like(x), x).copy(), tpl.mul(np.empty_like(x), x, add_b=False).copy()) for x in xs]
        assert tpl.engine.stats()["pipe_calls"] == 0
        for bands in (2, 5, 16):
            _pipeline_env(FVM_PIPE_MIN_NODES=0, FVM_PIPE_FORCE=1, FVM_PIPE_BANDS=bands)
            tpl = build_template(c, "gpu", tile_triangles=128)
            for x, (yb, y0) in zip(xs, refs):
                assert np.array_equal(tpl.mul(np.full_like(x, np.nan), x), yb)
                assert np.array_equal(tpl.mul(np.full_like(x, np.nan), x, add_b=False), y0)
            st = tpl.engine.stats()
            assert st["pipe_calls"] == 4 and st["pipe_bands"] == bands
        x, y = xs[1].copy(), np.empty_like(xs[1])
        with G.pinned(x, y):  # zero-copy bands on page-locked vectors
            _pipeline_env(FVM_PIPE_MIN_NODES=0, FVM_PIPE_FORCE=1, FVM_PIPE_BANDS=4, FVM_PIPE_ZC=3)
            tpl = build_template(c, "gpu", tile_triangles=128)
            y[...] = np.nan
            assert np.array_equal(tpl.mul(y, x), refs[1][0])
            y[...] = np.nan
            assert np.array_equal(tpl.mul(y, x, add_b=False), refs[1][1])
    finally:
        _pipeline_env()


def test_saveat_is_validated():
    """saveat times that the stepper can never hit (unsorted, outside the time span, off the fixed-step grid) are
    rejected up front instead of blocking every later save and returning unwritten rows."""
    from fvm_b200 import _lib as L
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 20, 20, single_boundary=True))
    ic = np.where(pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    tpl = G.DiffusionEquation(pair.gmesh, gBC, diffusion_function=1 / 9, initial_condition=ic, final_time=0.1)
    ok = G.solve(tpl, G.Tsit5(0.01), saveat=[0.0, 0.05, 0.1])
    assert len(ok.u) == 3 and all(np.isfinite(r).all() for r in ok.u)
    for bad in ([0.05, 0.02], [-0.1, 0.05], [0.05, 0.2], [0.055]):
        with pytest.raises(L.FVMCudaError, match="saveat"):
            G.solve(tpl, G.Tsit5(0.01), saveat=bad)
    for bad in ([0.05, 0.02], [0.05, 0.2]):  # the adaptive stepper stops at any time inside the span, in order
        with pytest.raises(L.FVMCudaError, match="saveat"):
            G.solve(tpl, G.Tsit5(), saveat=bad)
    assert len(G.solve(tpl, G.Tsit5(), saveat=[0.013, 0.0777]).u) == 2
    tpl.engine.close()


def test_fused_pcg_matches_unfused(monkeypatch):
    """PCG on one GPU takes p.(A p) from the SpMV kernels' epilogue (one partial per CTA, fixed-order final sum) instead
    of a separate dot kernel, and alpha / beta from the kernels that finish the sums.  The unfused iteration
    (FVM_NO_FUSE=1) is the same arithmetic up to summation order: same solution, iteration counts within 2."""
    gtri = _split_loop(delaunay_mesh(4000, 5, extra_points=3, jitter=0.3), k=2)  # interface rows, tail rows, points that are not vertices
    pair = Pair(gtri)
    gBC = G.BoundaryConditions(pair.gmesh, (G.Const(1.0), G.Const(0.0)), (G.Dirichlet, G.Dirichlet))

    def run_pcg():
        tpl = G.MeanExitTimeProblem(pair.gmesh, gBC, diffusion_function=1e-3, tile_triangles=256)
        out = G.solve(tpl, G.KrylovJacobi("pcg", rtol=1e-12))
        tpl.engine.close()
        return out

    fused_cg = run_pcg()
    monkeypatch.setenv("FVM_NO_FUSE", "1")
    plain_cg = run_pcg()
    assert fused_cg.relres <= 1e-11 and plain_cg.relres <= 1e-11
    assert abs(fused_cg.iters - plain_cg.iters) <= 2
    assert rel_err(fused_cg.u, plain_cg.u) <= 1e-9
