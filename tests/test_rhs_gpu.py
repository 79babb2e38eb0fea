"""GPU parity of the right-hand side: libfvmcuda (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Bar (BASELINE.json): one RHS evaluation within 1e-12 relative error (fp64,
only the summation order differs); geometry bit-exact; integer connectivity bit-exact."""
import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import RTOL_RHS, Pair, delaunay_mesh, rel_err

pytestmark = pytest.mark.gpu


def run_both(gp, op, u, t=0.0, **kw):
    p = G.get_cuda_parameters(gp, **kw)
    du = G.fvm_eqs(np.zeros_like(u), u, p, t)
    ref = O.fvm_eqs_vec(np.zeros_like(u), u, op, t)
    return du, ref, p


@pytest.mark.parametrize("geometry_mode", [0, 1])
@pytest.mark.parametrize("tile", [64, 256, 1024])
def test_readme_diffusion_50x50(geometry_mode, tile):
    """BASELINE config 1: triangulate_rectangle 50x50 on [0,2]^2, Dirichlet u=0, D=1/9."""
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True))
    ic = np.where(pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    gp, op = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), ic=ic, final_time=0.5)
    for u in (ic, 50 * np.random.default_rng(20240517).random(len(ic))):
        du, ref, p = run_both(gp, op, u, tile_triangles=tile, geometry_mode=geometry_mode)
        assert rel_err(du, ref) <= RTOL_RHS
        # loop oracle == vectorised oracle (bitwise) on this config
        assert np.array_equal(ref, O.fvm_eqs(np.zeros_like(u), u, op, 0.0))
        st = p.engine.stats()
        assert st["n_vertices"] == 2500 and st["n_dirichlet"] == 196 and st["n_live_boundary_edges"] == 0


def test_geometry_bit_exact_and_connectivity():
    """FVMGeometry (geometry.jl:99-169): s, cv-edge midpoints, normals, lengths bit-exact against the
    oracle; cv volumes to the ulp level (summation order); permutations are bijections."""
    for gtri in (G.triangulate_rectangle(0, 2, 0, 2, 37, 23, single_boundary=False), delaunay_mesh(700, 1, extra_points=5)):
        pair = Pair(gtri)
        gp, _ = pair.problem((G.Const(0.0),) * len(gtri.boundary_sections), (G.Neumann,) * len(gtri.boundary_sections),
                             G.ConstantDiffusion(1.0))
        p = G.get_cuda_parameters(gp, tile_triangles=128)
        geo = p.engine.geometry()
        om = pair.omesh
        assert np.array_equal(geo["s"], om.s)
        assert np.array_equal(geo["mid"], om.mid)
        assert np.array_equal(geo["nrm"], om.nrm)
        assert np.array_equal(geo["len"], om.len)
        assert rel_err(geo["cv_volumes"], om.cv_volumes) <= 4e-16
        node_perm, tri_perm = p.engine.permutation()
        assert np.array_equal(np.sort(node_perm), np.arange(gtri.num_points))
        assert np.array_equal(np.sort(tri_perm), np.arange(gtri.num_triangles))
        # FVMGeometry read-back without a problem
        assert rel_err(pair.gmesh.cv_volumes, geo["cv_volumes"]) <= 4e-16  # tiling changes the summation order


def test_convection_all_neumann_robin():
    """test/test_functions.jl:339-374 (heat convection): four Neumann sections, constant flux at the
    bottom, Robin (affine in u) at the top."""
    L, k, T0, Tinf, alpha, q, h = 1.0, 237.0, 10.0, 10.0, 80.0e-6, 10.0, 25.0
    pair = Pair(G.triangulate_rectangle(0, L, 0, L, 61, 47, single_boundary=False))
    specs = (G.Const(-alpha * q / k), G.Const(0.0), G.AffineU(-alpha * h / k * Tinf, alpha * h / k), G.Const(0.0))
    gp, op = pair.problem(specs, (G.Neumann,) * 4, G.ConstantDiffusion(alpha))
    u = T0 + np.random.default_rng(7).random(pair.gtri.num_points)
    for mode in (0, 1):
        du, ref, p = run_both(gp, op, u, tile_triangles=256, geometry_mode=mode)
        assert rel_err(du, ref) <= RTOL_RHS
        assert p.engine.stats()["n_live_boundary_edges"] == 2 * 60 + 2 * 46


def _split_loop(gtri, k=4):
    loop = gtri.boundary_sections[0]
    n = len(loop) - 1
    cuts = [round(i * n / k) for i in range(k + 1)]
    secs = [loop[cuts[i]:cuts[i + 1] + 1] for i in range(k)]
    return G.Triangulation(gtri.points, gtri.triangles, secs)


@pytest.mark.parametrize("geometry_mode", [0, 1])
def test_unstructured_mixed_conditions_power_diffusion(geometry_mode):
    """Delaunay mesh with points that are not vertices; Dirichlet, Dudt, Neumann and Constrained
    sections; internal Dirichlet and Dudt nodes; porous-medium diffusion D0*u^(m-1) with a
    logistic source (docs tutorials porous_medium_equation.jl:50, porous_fisher_equation...jl:60-61)."""
    gtri = _split_loop(delaunay_mesh(1500, 3, extra_points=7))
    pair = Pair(gtri)
    specs = (G.Const(0.25), G.AffineU(0.1, -0.5), G.LinearXY(0.3, 0.2, -0.1), G.Const(0.0))
    types = (G.Dirichlet, G.Dudt, G.Neumann, G.Constrained)
    internal = ((G.Const(0.7), G.AffineU(0.0, 1.0)), {200: 0, 201: 0}, {300: 1, 200: 1})
    rng = np.random.default_rng(11)
    u = 0.2 + rng.random(gtri.num_points)
    for flux in (G.PowerDiffusion(0.3, 2.0), G.PowerDiffusion(0.3, 2.5, use_abs=True), G.PowerDiffusion(0.3, 1.0)):
        gp, op = pair.problem(specs, types, flux, source=G.LogisticSource(1.3), internal=internal)
        du, ref, p = run_both(gp, op, u, t=0.3, tile_triangles=128, geometry_mode=geometry_mode)
        assert rel_err(du, ref) <= RTOL_RHS
        assert np.all(du[-7:] == 0.0)  # points that are not vertices
        # loop oracle agrees with the vectorised oracle
        assert rel_err(ref, O.fvm_eqs(np.zeros_like(u), u, op, 0.3)) <= 1e-15


def test_tabulated_diffusion_and_source():
    """(x,y)-only D and S are tabulated per cv-edge / node at setup (north_star b)."""
    gtri = delaunay_mesh(900, 5)
    pair = Pair(gtri)
    Dfn = lambda x, y: 1.0 + 0.5 * np.sin(3 * x) * np.cos(2 * y)
    Sfn = lambda x, y: np.exp(-x) * y
    gp, op = pair.problem(G.Const(0.0), G.Neumann, G.TabulatedDiffusion(Dfn), source=G.TabulatedSource(Sfn))
    u = np.random.default_rng(5).random(gtri.num_points)
    du, ref, _ = run_both(gp, op, u, tile_triangles=64)
    assert rel_err(du, ref) <= RTOL_RHS


def test_advection_diffusion_and_linear_source():
    """docs tutorial piecewise_linear_and_natural_neighbour...jl:80-88: q = (nu u - D u_x, -D u_y)."""
    pair = Pair(G.triangulate_rectangle(-1, 1, -0.5, 0.5, 48, 31, single_boundary=True))
    gp, op = pair.problem(G.Const(0.0), G.Dirichlet, G.AdvectionDiffusionFlux(0.02, 0.05, 0.0), source=G.LinearSource(-0.2, 0.1))
    P = pair.gtri.points
    u = np.exp(-(P[:, 0] ** 2 + P[:, 1] ** 2) / 0.01) / (0.01 * np.pi)
    for mode in (0, 1):
        du, ref, _ = run_both(gp, op, u, geometry_mode=mode, tile_triangles=256)
        assert rel_err(du, ref) <= RTOL_RHS


@pytest.mark.parametrize("geometry_mode", [0, 1])
def test_system_gray_scott_and_keller_segel(geometry_mode):
    """FVMSystem (problem.jl:233-279): species-interleaved state, per-species conditions, fluxes that
    see every species (Keller-Segel, src/FiniteVolumeMethod.jl:92-138) and coupled sources."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 40, 33, single_boundary=True))
    N = pair.gtri.num_points
    rng = np.random.default_rng(13)
    U = np.ascontiguousarray(np.stack([0.5 + 0.5 * rng.random(N), 0.25 * rng.random(N)], axis=1))
    # Gray-Scott: zero-flux Neumann, constant per-species diffusion
    src = G.GrayScottSource(0.04, 0.1)
    g1, o1 = pair.problem(G.Const(0.0), G.Neumann, G.ConstantDiffusion(2e-5), source=src, var=0, ic=U[:, 0])
    g2, o2 = pair.problem(G.Const(0.0), G.Neumann, G.ConstantDiffusion(1e-5), source=src, var=1, ic=U[:, 1])
    du, ref, _ = run_both(G.FVMSystem(g1, g2), O.FVMSystem(o1, o2), U, tile_triangles=256, geometry_mode=geometry_mode)
    assert du.shape == (N, 2) and rel_err(du, ref) <= RTOL_RHS
    # Keller-Segel with a Dirichlet condition on species 0 only and a Robin Neumann edge on species 1
    ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    g1, o1 = pair.problem(G.Const(0.3), G.Dirichlet, ks, source=kss, var=0, ic=U[:, 0])
    g2, o2 = pair.problem(G.AffineU(0.01, -0.2), G.Neumann, ks, source=kss, var=1, ic=U[:, 1])
    du, ref, p = run_both(G.FVMSystem(g1, g2), O.FVMSystem(o1, o2), U, t=1.5, tile_triangles=128, geometry_mode=geometry_mode)
    assert rel_err(du[:, 0], ref[:, 0]) <= RTOL_RHS and rel_err(du[:, 1], ref[:, 1]) <= RTOL_RHS
    # system of two identical diffusion problems reproduces the scalar problem (test/equations.jl:106-121)
    g1, o1 = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), var=0, ic=U[:, 0])
    g2, o2 = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), var=1, ic=U[:, 0])
    gs, _ = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), ic=U[:, 0])
    UU = np.ascontiguousarray(np.stack([U[:, 0], U[:, 0]], axis=1))
    dsys = G.fvm_eqs(np.zeros_like(UU), UU, G.get_cuda_parameters(G.FVMSystem(g1, g2)), 0.0)
    dsc = G.fvm_eqs(np.zeros(N), np.ascontiguousarray(U[:, 0]), G.get_cuda_parameters(gs), 0.0)
    assert rel_err(dsys[:, 0], dsc) <= 1e-14 and np.array_equal(dsys[:, 0], dsys[:, 1])


def test_dirichlet_callback_time_dependent():
    """update_dirichlet_nodes! (dirichlet.jl:78-86) with the annulus tutorial's 50(1-exp(-t/2))."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 20, 20, single_boundary=False))
    specs = (G.ExpSaturation(50.0, 2.0), G.Const(1.0), G.LinearXY(0.0, 1.0, 2.0), G.AffineU(0.5, 0.5))
    gp, op = pair.problem(specs, (G.Dirichlet,) * 4, G.ConstantDiffusion(1.0))
    u = np.random.default_rng(3).random(400)
    p = G.get_cuda_parameters(gp)
    got = G.update_dirichlet_nodes(u.copy(), 0.7, p)
    ref = O.update_dirichlet_nodes(u.copy(), 0.7, op)
    # corners belong to two sections: the later section wins in both implementations
    assert rel_err(got, ref) <= 1e-15


def test_determinism_and_midsize_lattice():
    """512x512 lattice (262k nodes): parity at a size the vectorised oracle finishes in seconds,
    and bitwise run-to-run determinism (no atomics)."""
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 512, 512, single_boundary=True))
    N = pair.gtri.num_points
    u = 50 * np.random.default_rng(20240517).random(N)
    gp, op = pair.problem(G.Const(0.0), G.Dirichlet, G.PowerDiffusion(1 / 9, 1.0))
    du, ref, p = run_both(gp, op, u)
    assert rel_err(du, ref) <= RTOL_RHS
    du2 = G.fvm_eqs(np.zeros_like(u), u, p, 0.0)
    assert np.array_equal(du, du2)
    # recompute-geometry kernel at the same size: Delta and the s7..s9 numerators cancel at eps*|x|^2/area,
    # so they must be rounded exactly like the reference (a contracted FMA shows up at 1e-10 here)
    du_r = G.fvm_eqs(np.zeros_like(u), u, G.get_cuda_parameters(gp, geometry_mode=1), 0.0)
    assert rel_err(du_r, ref) <= RTOL_RHS
    # linearity of the constant-coefficient operator: F(a u + b w) = a F(u) + b F(w)
    w = np.random.default_rng(1).random(N)
    for g in (u, w):
        g[p.prob.conditions.node_kind != 0] = 0.0
    Fu = G.fvm_eqs(np.zeros_like(u), u, p, 0.0)
    Fw = G.fvm_eqs(np.zeros_like(u), w, p, 0.0)
    Fc = G.fvm_eqs(np.zeros_like(u), 2.0 * u - 3.0 * w, p, 0.0)
    assert rel_err(Fc, 2.0 * Fu - 3.0 * Fw) <= 1e-12


def test_sparse_jacobian_matches_finite_differences():
    """d(fvm_eqs!)/du on the device (exact duals) vs central differences of the ORACLE right-hand side,
    and its pattern vs jacobian_sparsity (solve.jl:56-131), scalar and FVMSystem."""
    gtri = _split_loop(delaunay_mesh(600, 31, extra_points=2))
    pair = Pair(gtri)
    N = gtri.num_points
    rng = np.random.default_rng(17)
    specs = (G.Const(0.25), G.AffineU(0.1, -0.5), G.AffineU(0.3, 0.7), G.Const(0.0))
    types = (G.Dirichlet, G.Dudt, G.Neumann, G.Constrained)
    cases = [(G.PowerDiffusion(0.3, 2.5, use_abs=True), G.LogisticSource(1.3)), (G.AdvectionDiffusionFlux(0.02, 0.5, -0.3), G.LinearSource(-0.2, 0.1)),
             (G.TabulatedDiffusion(lambda x, y: 1.0 + x * y), None)]
    for flux, src in cases:
        gp, op = pair.problem(specs, types, flux, source=src)
        p = G.get_cuda_parameters(gp, tile_triangles=128)
        u = 0.5 + rng.random(N)
        J = G.jacobian(u, p, 0.2)
        r, c = O.jacobian_sparsity(pair.otri)
        P = G.jacobian_sparsity(p)
        assert P.nnz == len(r) and (P[r, c] == 1).all()
        v = rng.standard_normal(N)
        eps = 1e-6
        fd = (O.fvm_eqs_vec(np.zeros(N), u + eps * v, op, 0.2) - O.fvm_eqs_vec(np.zeros(N), u - eps * v, op, 0.2)) / (2 * eps)
        assert rel_err(J @ v, fd) <= 1e-6
    # FVMSystem: Keller-Segel flux + sources, Robin Neumann edge on species 1, Dirichlet on species 0
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 17, 13, single_boundary=True))
    N = pair.gtri.num_points
    U = np.ascontiguousarray(np.stack([0.5 + 0.5 * rng.random(N), 0.25 * rng.random(N)], axis=1))
    ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    g1, o1 = pair.problem(G.Const(0.3), G.Dirichlet, ks, source=kss, var=0, ic=U[:, 0])
    g2, o2 = pair.problem(G.AffineU(0.01, -0.2), G.Neumann, ks, source=kss, var=1, ic=U[:, 1])
    gs, os_ = G.FVMSystem(g1, g2), O.FVMSystem(o1, o2)
    p = G.get_cuda_parameters(gs, tile_triangles=128)
    J = G.jacobian(U, p, 0.0)
    r, c = O.jacobian_sparsity(pair.otri, 2)
    assert J.shape == (2 * N, 2 * N) and G.jacobian_sparsity(p).nnz == len(r)
    V = rng.standard_normal((N, 2))
    eps = 1e-6
    fd = (O.fvm_eqs_vec(np.zeros_like(U), U + eps * V, os_, 0.0) - O.fvm_eqs_vec(np.zeros_like(U), U - eps * V, os_, 0.0)) / (2 * eps)
    assert rel_err((J @ V.ravel()).reshape(N, 2), fd) <= 1e-6


def test_three_species_and_abi_error_codes():
    """neq = 3 (FVMSystem{3}) reproduces three scalar problems; the ABI's error behaviour: call order,
    unknown registry ids (the analogue of InvalidFluxError, problem.jl:291-316), bad sizes."""
    import ctypes as C
    from fvm_b200 import _lib as L
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 23, 19, single_boundary=True))
    N = pair.gtri.num_points
    rng = np.random.default_rng(23)
    U = np.ascontiguousarray(rng.random((N, 3)))
    Ds = (0.3, 0.7, 1.1)
    gps, scal = [], []
    for v, D in enumerate(Ds):
        g, _ = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(D), source=G.LinearSource(-0.1 * (v + 1), 0.2), var=v, ic=U[:, v])
        gps.append(g)
        gs, os_ = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(D), source=G.LinearSource(-0.1 * (v + 1), 0.2), ic=U[:, v])
        scal.append(O.fvm_eqs_vec(np.zeros(N), np.ascontiguousarray(U[:, v]), os_, 0.0))
    p = G.get_cuda_parameters(G.FVMSystem(*gps))
    dU = G.fvm_eqs(np.zeros_like(U), U, p, 0.0)
    for v in range(3):
        assert rel_err(dU[:, v], scal[v]) <= RTOL_RHS
    # ---- error codes through the raw ABI ----
    lib = L.lib()
    tri = pair.gtri
    h = L.H()
    pts, tr = L.f64(tri.points), L.i32(tri.triangles)
    assert lib.fvm_create(L.dp(pts), N, L.ip(tr), tri.num_triangles, 0, 1, 0, C.byref(h)) == L.OK
    buf = np.zeros(N)
    assert lib.fvm_rhs(h, 0.0, buf.ctypes.data, buf.ctypes.data, 0) == L.ERR_STATE      # rhs before finalize
    assert b"finalize" in lib.fvm_last_error(h)
    one = L.f64([1.0])
    assert lib.fvm_set_flux(h, 99, L.dp(one), 1) == L.ERR_UNSUPPORTED                   # closure outside the registry
    assert b"registry" in lib.fvm_last_error(h)
    assert lib.fvm_set_flux(h, G.FLUX_KELLER_SEGEL, L.dp(L.f64([4.0, 1.0])), 2) == L.ERR_ARG  # needs neq == 2
    assert lib.fvm_set_flux(h, G.FLUX_DIFF_CONST, L.dp(L.f64([1.0, 2.0])), 2) == L.ERR_ARG    # wrong parameter count
    assert lib.fvm_finalize(h, 100, 0) == L.ERR_ARG                                     # tile size not a multiple of 64
    assert lib.fvm_finalize(h, 0, 0) == L.OK
    assert lib.fvm_finalize(h, 0, 0) == L.ERR_STATE                                     # already finalized
    assert lib.fvm_spmv(h, buf.ctypes.data, buf.ctypes.data, 1, 0) == L.ERR_STATE       # spmv before assemble
    bad = L.i32([[0, 1, N + 5]])
    h2 = L.H()
    assert lib.fvm_create(L.dp(pts), N, L.ip(bad), 1, 0, 1, 0, C.byref(h2)) == L.ERR_ARG  # vertex out of range
    assert lib.fvm_create(L.dp(pts), N, L.ip(tr), tri.num_triangles, 0, 9, 0, C.byref(h2)) == L.ERR_ARG
    lib.fvm_destroy(h)


def test_pl_interpolate_and_compute_flux():
    """pl_interpolate (utils.jl:23-27) and compute_flux on every edge (problem.jl:458-487;
    test/test_functions.jl:659-747 checks it on every edge too), scalar and FVMSystem."""
    gtri = delaunay_mesh(400, 41)
    pair = Pair(gtri)
    N = gtri.num_points
    rng = np.random.default_rng(29)
    u = 0.3 + rng.random(N)
    gp, op = pair.problem(G.Const(0.0), G.Neumann, G.PowerDiffusion(0.3, 2.0))
    p = G.get_cuda_parameters(gp, tile_triangles=64)
    T = gtri.triangles
    E = np.unique(np.sort(np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]), axis=1), axis=0)
    E = np.concatenate([E, E[:, ::-1]])  # both orientations
    got = G.compute_flux(p, E[:, 0], E[:, 1], u, 0.4)
    ref = np.array([O.compute_flux(op, int(a), int(b), u, 0.4) for a, b in E])
    assert rel_err(got, ref) <= 1e-12
    tq = rng.integers(0, len(T), 300)
    w = rng.dirichlet((1, 1, 1), 300)
    pts = np.einsum("nk,nkd->nd", w, gtri.points[T[tq]])
    got = G.pl_interpolate(p, tq, u, pts[:, 0], pts[:, 1])
    ref = np.array([O.pl_interpolate(op, T[t], u, x, y) for t, (x, y) in zip(tq, pts)])
    assert rel_err(got, ref) <= 1e-12
    assert rel_err(got, np.einsum("nk,nk->n", w, u[T[tq]])) <= 1e-10  # barycentric interpolation
    # system
    U = np.ascontiguousarray(np.stack([u, 0.5 * u[::-1]], axis=1))
    ks = G.KellerSegelFlux(4.0, 1.0)
    g1, o1 = pair.problem(G.Const(0.0), G.Neumann, ks, var=0, ic=U[:, 0])
    g2, o2 = pair.problem(G.Const(0.0), G.Neumann, ks, var=1, ic=U[:, 1])
    ps = G.get_cuda_parameters(G.FVMSystem(g1, g2), tile_triangles=64)
    got = G.compute_flux(ps, E[:50, 0], E[:50, 1], U, 0.0)
    ref = np.array([O.compute_flux(O.FVMSystem(o1, o2), int(a), int(b), U, 0.0) for a, b in E[:50]])
    assert got.shape == (50, 2) and rel_err(got, ref) <= 1e-12


def test_abi_index_base_one_and_native_entry_points():
    """Julia passes 1-based triangles / edges (index_base = 1): same result as 0-based.  Device-pointer
    entry points: fvm_to_native -> fvm_rhs_native -> fvm_from_native equals fvm_rhs (caller order)."""
    import ctypes as C
    import torch
    from fvm_b200 import _lib as L
    lib = L.lib()
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 1, 31, 17, single_boundary=False))
    tri = pair.gtri
    N = tri.num_points
    gp, op = pair.problem((G.Const(0.1), G.Const(0.0), G.AffineU(0.2, -0.3), G.Const(0.0)),
                          (G.Neumann, G.Dirichlet, G.Neumann, G.Dudt), G.ConstantDiffusion(0.4))
    u = np.random.default_rng(31).random(N)
    ref = G.fvm_eqs(np.zeros(N), u, G.get_cuda_parameters(gp), 0.0)
    c = gp.conditions
    h = L.H()
    pts, tr1 = L.f64(tri.points), L.i32(tri.triangles + 1)
    assert lib.fvm_create(L.dp(pts), N, L.ip(tr1), tri.num_triangles, 1, 1, 0, C.byref(h)) == L.OK
    uv1 = L.i32(c.boundary_edges + 1)
    L.check(h, lib.fvm_set_boundary_edges(h, L.ip(uv1), len(uv1)))
    ek, ef = np.ascontiguousarray(c.edge_kind), L.i32(c.edge_fidx)
    nk, nf = np.ascontiguousarray(c.node_kind), L.i32(c.node_fidx)
    L.check(h, lib.fvm_set_edge_conditions(h, 0, L.bp(ek), L.ip(ef)))
    L.check(h, lib.fvm_set_node_conditions(h, 0, L.bp(nk), L.ip(nf)))
    for fidx, spec in enumerate(c.functions):
        fid, params = G.functors.cond_spec(spec)
        pp = L.f64(params)
        L.check(h, lib.fvm_set_condition_fn(h, 0, fidx, fid, L.dp(pp), len(pp)))
    d = L.f64([0.4])
    L.check(h, lib.fvm_set_flux(h, G.FLUX_DIFF_CONST, L.dp(d), 1))
    L.check(h, lib.fvm_finalize(h, 128, 0))
    du = np.zeros(N)
    L.check(h, lib.fvm_rhs(h, 0.0, u.ctypes.data, du.ctypes.data, 0))
    assert rel_err(du, ref) <= 1e-13  # another tile size: only the summation order differs
    ref = du
    # device pointers, native order
    ud = torch.from_numpy(u).cuda()
    un, dn, dc = torch.empty_like(ud), torch.empty_like(ud), torch.empty_like(ud)
    L.check(h, lib.fvm_to_native(h, ud.data_ptr(), un.data_ptr()))
    L.check(h, lib.fvm_rhs_native(h, 0.0, un.data_ptr(), dn.data_ptr()))
    L.check(h, lib.fvm_from_native(h, dn.data_ptr(), dc.data_ptr()))
    L.check(h, lib.fvm_stream_synchronize(h))
    assert np.array_equal(dc.cpu().numpy(), ref)
    dd = torch.empty_like(ud)
    L.check(h, lib.fvm_rhs(h, 0.0, ud.data_ptr(), dd.data_ptr(), 1))
    assert np.array_equal(dd.cpu().numpy(), ref)
    node_perm = np.empty(N, np.int32)
    L.check(h, lib.fvm_get_permutation(h, L.ip(node_perm), None))
    assert np.array_equal(un.cpu().numpy(), u[node_perm])
    lib.fvm_destroy(h)


def test_mesh_from_wire_container_gives_identical_rhs(tmp_path):
    """fvm_create_from_wire (FVMWIRE container written 1-based, as the Julia side does) == fvm_create on the
    in-memory arrays: bitwise the same RHS and geometry; the saved solution round-trips."""
    tri = delaunay_mesh(400, seed=5, jitter=0.3)
    pair = Pair(tri)
    gp, op = pair.problem(G.Const(0.3), G.Neumann, G.PowerDiffusion(0.7, 2.0))
    u = 0.5 + np.random.default_rng(5).random(tri.num_points)
    path = str(tmp_path / "mesh.fvmw")
    G.save_mesh(path, tri)
    p_mem = G.get_cuda_parameters(gp, tile_triangles=128)
    p_file = G.get_cuda_parameters(gp, tile_triangles=128, mesh_file=path)
    du_mem = G.fvm_eqs(np.zeros_like(u), u, p_mem, 0.0)
    du_file = G.fvm_eqs(np.zeros_like(u), u, p_file, 0.0)
    assert np.array_equal(du_mem, du_file)
    assert rel_err(du_file, O.fvm_eqs_vec(np.zeros_like(u), u, op, 0.0)) <= RTOL_RHS
    for a, b in zip(p_mem.engine.geometry(), p_file.engine.geometry()):
        assert np.array_equal(a, b)
    sol = G.Solution(np.stack([u, du_file]), t=np.array([0.0, 1.0]))
    G.save_solution(str(tmp_path / "sol.fvmw"), sol)
    assert np.array_equal(G.load_solution(str(tmp_path / "sol.fvmw")).u, sol.u)


def _pipeline_env(**kw):
    import os
    keys = ("FVM_NO_PIPELINE", "FVM_PIPE_MIN_NODES", "FVM_PIPE_FORCE", "FVM_PIPE_BANDS", "FVM_PIPE_ZC", "FVM_PIPE_ZC_CTAS")
    for k in keys:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


@pytest.mark.parametrize("case", ["readme", "convection", "unstructured", "system"])
def test_host_buffer_pipeline_is_bit_identical(case):
    """fvm_rhs with host buffers runs as a banded copy-in / compute / copy-out pipeline on large meshes
    (fvm_pipe.cu).  Forced on here for small ones: every band count must reproduce the unpipelined call bit for
    bit (same kernels, same summation order), repeatedly on one handle and interleaved with device-pointer calls."""
    import torch
    from tests.golden_cases import CASES
    name = {"readme": "readme_50x50", "convection": "convection_robin_24", "unstructured": "unstructured_porous",
            "system": "keller_segel_16"}[case]
    c = CASES[name](None)
    try:
        _pipeline_env(FVM_NO_PIPELINE=1)
        p0 = G.get_cuda_parameters(c.gp, tile_triangles=64)
        ref = G.fvm_eqs(np.zeros_like(c.u), c.u, p0, c.t)
        assert p0.engine.stats()["pipe_calls"] == 0
        u2 = c.u[::-1].copy() if c.u.ndim == 1 else c.u[::-1, :].copy()
        ref2 = G.fvm_eqs(np.zeros_like(u2), u2, p0, c.t)
        for bands in (2, 3, 7, 16):
            _pipeline_env(FVM_PIPE_MIN_NODES=0, FVM_PIPE_FORCE=1, FVM_PIPE_BANDS=bands)
            p = G.get_cuda_parameters(c.gp, tile_triangles=64)
            du = G.fvm_eqs(np.full_like(c.u, np.nan), c.u, p, c.t)
            st = p.engine.stats()
            assert st["pipe_calls"] == 1 and st["pipe_bands"] == bands
            assert np.array_equal(du, ref)
            assert np.array_equal(G.fvm_eqs(np.full_like(u2, np.nan), u2, p, c.t), ref2)
            ud = torch.from_numpy(c.u).cuda()
            dd = torch.empty_like(ud)
            p.engine.rhs_device(dd.data_ptr(), ud.data_ptr(), c.t)
            assert np.array_equal(dd.cpu().numpy(), ref)
            assert np.array_equal(G.fvm_eqs(np.full_like(c.u, np.nan), c.u, p, c.t), ref)
            assert p.engine.stats()["pipe_calls"] == 3
        # page-locked caller arrays: the bands cross PCIe inside the renumbering kernels (zero-copy, FVM_PIPE_ZC bit 0 =
        # copy-in, bit 1 = copy-out); every combination must give the same bits, also with very few CTAs
        uz, dz = c.u.copy(), np.full_like(c.u, np.nan)
        with G.pinned(uz, dz):
            for zc, ctas in ((0, 64), (1, 64), (2, 3), (3, 1), (3, 64)):
                _pipeline_env(FVM_PIPE_MIN_NODES=0, FVM_PIPE_FORCE=1, FVM_PIPE_BANDS=5, FVM_PIPE_ZC=zc, FVM_PIPE_ZC_CTAS=ctas)
                p = G.get_cuda_parameters(c.gp, tile_triangles=64)
                dz[...] = np.nan
                assert np.array_equal(G.fvm_eqs(dz, uz, p, c.t), ref)
                uz[...] = u2
                dz[...] = np.nan
                assert np.array_equal(G.fvm_eqs(dz, uz, p, c.t), ref2)  # the host wrote new values: the kernel must see them
                uz[...] = c.u
                assert p.engine.stats()["pipe_calls"] == 2
    finally:
        _pipeline_env()


def test_host_register_round_trip():
    """fvm_host_register / fvm_host_unregister (G.pinned): page-locked caller arrays give the same result; a
    second registration of the same array is accepted, a bad pointer is an error."""
    from fvm_b200 import _lib as L
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True))
    gp, op = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9))
    p = G.get_cuda_parameters(gp)
    u = np.random.default_rng(2).random(2500)
    ref = G.fvm_eqs(np.zeros_like(u), u, p, 0.0)
    du = np.zeros_like(u)
    with G.pinned(u, du):
        assert L.lib().fvm_host_register(u.ctypes.data, u.nbytes) == L.OK  # already registered: accepted
        assert np.array_equal(G.fvm_eqs(du, u, p, 0.0), ref)
    assert L.lib().fvm_host_unregister(u.ctypes.data) != L.OK  # no longer registered
    assert L.lib().fvm_host_register(None, 8) == L.ERR_ARG


def test_recompute_geometry_is_bit_identical_to_reference_arithmetic():
    """geometry_mode 1 recomputes geometry.jl:107-161 per triangle with contracted FMAs, a Newton reciprocal and
    Markstein-corrected quotients instead of IEEE divisions.  The variant the u-dependent fluxes run (they amplify an
    ulp of s by |x|/h) must equal the individually rounded reference arithmetic in every bit; the variant of the
    alpha/beta-only fluxes in every bit of s7..s9, cv-edge midpoints and vectors, and to one ulp in s1..s6 (lattices at h = 1.3e-3 where Delta cancels to 1e-10, a stretched lattice far
    from the origin, a jittered Delaunay mesh)."""
    import ctypes as C
    from fvm_b200 import _lib as L
    meshes = [G.triangulate_rectangle(0, 2, 0, 2, 1500, 1500, single_boundary=True),
              G.triangulate_rectangle(-3, 40, 1e3, 1e3 + 1, 300, 700, single_boundary=True),
              delaunay_mesh(60000, 5, jitter=0.35)]
    for tri in meshes:
        mesh = G.FVMGeometry(tri)
        prob = G.FVMProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann), diffusion_function=G.ConstantDiffusion(1.0),
                            initial_condition=np.zeros(tri.num_points), final_time=1.0)
        p = G.get_cuda_parameters(prob, geometry_mode=1)
        bad = C.c_int64(-1)
        L.check(p.engine.h, L.lib().fvm_check_recompute_geometry(p.engine.h, C.byref(bad)))
        assert bad.value == 0, "%d of %d triangles differ" % (bad.value, tri.num_triangles)
        p.engine.close()
