"""Parity at BASELINE.json's full sizes: DIRECT 1e-12 comparison of one fvm_eqs! at 4096^2 (configs 2 and 4, both
geometry modes) with the serial C oracle (oracle_fvm_eqs_general, bitwise the NumPy oracle on small meshes),
config 3 at 1024^2 against the SuperLU oracle, plus size-independent properties (conservation, linearity,
operator == RHS, determinism) and a 200k-point unstructured mesh against the vectorised oracle."""
import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import RTOL_RHS, Pair, delaunay_mesh, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    tri = G.triangulate_rectangle(0, 2, 0, 2, 4096, 4096, single_boundary=True)
    return tri, G.FVMGeometry(tri)


@pytest.mark.parametrize("flux,fm,fp", [(G.ConstantDiffusion(1 / 9), 0, [1 / 9]), (G.PowerDiffusion(1 / 9, 2.0), 2, [1 / 9, 2.0, 0.0])],
                         ids=["const_D", "u_dependent_D"])
def test_config2_rhs_4096_direct_parity(big, flux, fm, fp):
    """BASELINE config 2 mesh (16.7M nodes, 33.5M triangles), README conditions (Dirichlet u = 0): one fvm_eqs! of
    the streaming recompute kernel (the engine's default) and of the stored-geometry kernel against the serial
    C oracle, 1e-12 relative -- the bar of /root/reference/test/test_functions.jl:645-657 at full size."""
    from oracle import c_oracle
    tri, mesh = big
    N = tri.num_points
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    u = 50 * np.random.default_rng(20240517).random(N)
    bnd = np.zeros(N, bool)
    bnd[np.unique(tri.boundary_edges()[0])] = True
    ref = c_oracle.fvm_eqs_general(tri.points, tri.triangles, u, 1, fm, fp, dirichlet=bnd)
    prob = G.FVMProblem(mesh, BC, diffusion_function=flux, initial_condition=u, final_time=1.0)
    for mode in (1, 0):
        p = G.get_cuda_parameters(prob, geometry_mode=mode)
        du = G.fvm_eqs(np.empty(N), u, p, 0.0)
        assert rel_err(du, ref) <= RTOL_RHS, (mode, rel_err(du, ref))
        p.engine.close()


def test_config4_system_rhs_4096_direct_parity(big):
    """BASELINE config 4: 2-species Keller-Segel FVMSystem (src/FiniteVolumeMethod.jl:92-138), all-Neumann zero flux,
    4096^2, against the serial C oracle at 1e-12 per species."""
    from oracle import c_oracle
    tri, mesh = big
    N = tri.num_points
    rng = np.random.default_rng(20240517)
    U = np.ascontiguousarray(np.stack([0.01 * rng.random(N), 0.005 * rng.random(N)], axis=1))
    ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann)
    pu = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=U[:, 0], final_time=1.0)
    pv = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=U[:, 1], final_time=1.0)
    ref = c_oracle.fvm_eqs_general(tri.points, tri.triangles, U, 2, 4, [4.0, 1.0], 6, [0.1])
    p = G.get_cuda_parameters(G.FVMSystem(pu, pv))
    du = G.fvm_eqs(np.empty_like(U), U, p, 0.0)
    assert rel_err(du[:, 0], ref[:, 0]) <= RTOL_RHS and rel_err(du[:, 1], ref[:, 1]) <= RTOL_RHS
    p.engine.close()


def test_config3_mean_exit_time_1024_vs_superlu_oracle():
    """BASELINE config 3 at 1024^2 (1.05M unknowns): the assembled MeanExitTimeProblem matches the oracle's
    (bitwise hand assembly in the reference: docs/src/literate_wyos/mean_exit_time.jl:206) to 1e-12 ||A||, and the
    Jacobi-PCG solution agrees with the oracle's SuperLU solve to the solver tolerance."""
    n = 1024
    gtri = G.triangulate_rectangle(0, 2, 0, 2, n, n, single_boundary=True)
    pair = Pair(gtri)
    gBC = G.BoundaryConditions(pair.gmesh, G.Const(0.0), G.Dirichlet)
    oBC = O.BoundaryConditions(pair.omesh, (lambda x, y, t, u, p: 0.0,), (O.Dirichlet,))
    met = G.MeanExitTimeProblem(pair.gmesh, gBC, diffusion_function=1 / 9)
    ref = O.MeanExitTimeProblem(pair.omesh, oBC, diffusion_function=lambda x, y, p: 1 / 9, vectorised=True)
    assert abs(met.A - ref.A).max() <= 1e-12 * abs(ref.A).max()
    assert rel_err(met.b, ref.b) <= 1e-12
    sol = G.solve(met, G.KrylovJacobi("pcg", rtol=1e-12, maxiter=40000))
    uo = O.solve_steady(ref)
    assert rel_err(sol.u, uo) <= 1e-8
    met.engine.close()


def test_conservation_and_determinism_4096(big):
    """All-Neumann zero flux, no source: sum_i V_i du_i = 0 (every cv-edge flux enters two control
    volumes with opposite signs), for a u-dependent flux; bitwise run-to-run determinism."""
    tri, mesh = big
    N = tri.num_points
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann)
    u = 1.0 + np.random.default_rng(20240517).random(N)
    prob = G.FVMProblem(mesh, BC, diffusion_function=G.PowerDiffusion(1 / 9, 2.0), initial_condition=u, final_time=1.0)
    p = G.get_cuda_parameters(prob)
    du = G.fvm_eqs(np.empty(N), u, p, 0.0)
    V = p.engine.geometry_volumes() if hasattr(p.engine, "geometry_volumes") else None
    if V is None:
        V = np.empty(N)
        from fvm_b200 import _lib as L
        L.check(p.engine.h, L.lib().fvm_get_geometry(p.engine.h, L.dp(V), None, None, None, None))
    flux_sum = float(np.dot(V, du))
    scale = float(np.dot(V, np.abs(du)))
    assert abs(flux_sum) <= 1e-12 * scale
    assert abs(V.sum() - 4.0) <= 1e-12 * 4.0  # control volumes tile the domain
    assert np.array_equal(du, G.fvm_eqs(np.empty(N), u, p, 0.0))
    st = p.engine.stats()
    assert st["n_vertices"] == N and st["n_live_boundary_edges"] == 4 * 4095


def test_linearity_and_operator_equals_rhs_4096(big):
    """README problem at 4096^2: F is linear for constant D, and the DiffusionEquation template
    operator A u + b reproduces fvm_eqs!(u) (docs/src/literate_wyos/diffusion_equations.jl:436-438)."""
    tri, mesh = big
    N = tri.num_points
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    prob = G.FVMProblem(mesh, BC, diffusion_function=G.ConstantDiffusion(1 / 9), initial_condition=ic, final_time=0.5)
    p = G.get_cuda_parameters(prob)
    rng = np.random.default_rng(1)
    u, w = 50 * rng.random(N), rng.random(N)
    Fu = G.fvm_eqs(np.empty(N), u, p, 0.0)
    Fw = G.fvm_eqs(np.empty(N), w, p, 0.0)
    Fc = G.fvm_eqs(np.empty(N), 2.0 * u - 3.0 * w, p, 0.0)
    assert rel_err(Fc, 2.0 * Fu - 3.0 * Fw) <= 1e-12
    # host buffers of this size take the banded copy/compute pipeline (fvm_pipe.cu): bit-identical to the
    # device-pointer call, and on the reference's row-major numbering all but the last band leave early
    import torch
    ud = torch.from_numpy(u).cuda()
    dd = torch.empty_like(ud)
    p.engine.rhs_device(dd.data_ptr(), ud.data_ptr(), 0.0)
    assert np.array_equal(dd.cpu().numpy(), Fu)
    st = p.engine.stats()
    # 4 host-buffer calls so far: pipeline warm-up, pipeline timed, plain timed, then the faster of the two
    assert st["pipe_calls"] in (2, 3) and st["host_schedule_rhs"] in ("pipeline", "plain") and st["pipe_bands"] >= 2 and st["pipe_early_bands"] >= st["pipe_bands"] - 2
    del ud, dd
    # interior stencil D/h^2 [1,1,-4,1,1] on a smooth field: F(x^2 + y^2) = 4 D away from the boundary
    P = tri.points
    q = G.fvm_eqs(np.empty(N), P[:, 0] ** 2 + P[:, 1] ** 2, p, 0.0)
    inner = (np.arange(N) % 4096 > 0) & (np.arange(N) % 4096 < 4095) & (np.arange(N) >= 4096) & (np.arange(N) < N - 4096)
    assert np.abs(q[inner] - 4.0 / 9.0).max() <= 1e-5  # h^-2 ~ 4e6 amplifies rounding: 1e-16 * 4e6 * 50
    p.engine.close()
    tpl = G.DiffusionEquation(mesh, BC, diffusion_function=1 / 9, initial_condition=ic, final_time=0.5)
    Au = tpl.mul(np.empty(N), u)
    assert rel_err(Au, Fu) <= RTOL_RHS
    assert tpl.engine.stats()["nnz"] == 117407746  # N + 2E


def test_unstructured_200k_vs_oracle():
    """A 200k-node unstructured mesh (jittered lattice, Delaunay connectivity, node degrees 4..8: Hilbert
    tiling, gather lists and sliced-ELL slices with ragged rows) against the vectorised oracle.  (On
    uniformly random points the thin triangles make |du| ~ 1e10 and the comparison meaningless.)"""
    gtri = delaunay_mesh(200000, 77, jitter=0.35)
    pair = Pair(gtri)
    u = 0.2 + np.random.default_rng(3).random(gtri.num_points)
    for flux, src in ((G.PowerDiffusion(0.3, 2.0), G.LogisticSource(1.3)), (G.ConstantDiffusion(0.7), None)):
        gp, op = pair.problem(G.Const(0.0), G.Neumann, flux, source=src)
        for mode in (0, 1):
            p = G.get_cuda_parameters(gp, geometry_mode=mode)
            du = G.fvm_eqs(np.empty_like(u), u, p, 0.0)
            ref = O.fvm_eqs_vec(np.zeros_like(u), u, op, 0.0)
            assert rel_err(du, ref) <= RTOL_RHS
            p.engine.close()
