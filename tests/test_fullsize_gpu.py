"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot
finish 16.7M nodes in seconds): conservation, linearity, operator == RHS, determinism, and a
200k-point unstructured mesh against the vectorised oracle."""
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
    assert st["pipe_calls"] == 3 and st["pipe_bands"] >= 2 and st["pipe_early_bands"] >= st["pipe_bands"] - 2
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
