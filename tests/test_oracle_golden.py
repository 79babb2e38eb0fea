"""Pins the CPU oracle against the reference's own numeric known-answer tests
(SURVEY.md section 8c).  Every test cites the reference test it restates."""
import math
import os

import numpy as np
import pytest

from oracle import fvm_oracle as O


def isapprox(x, y, rtol=None, atol=0.0):
    """Julia's isapprox for arrays: norm(x-y) <= max(atol, rtol*max(norm(x),norm(y)))."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if rtol is None:
        rtol = math.sqrt(np.finfo(float).eps) if atol == 0 else 0.0
    return np.linalg.norm(x - y) <= max(atol, rtol * max(np.linalg.norm(x), np.linalg.norm(y)))


# ---- problems of /root/reference/test/test_functions.jl:304-374 ------------------------
def example_diffusion_problem():
    tri = O.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, 25, 25, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0 * u, O.Dirichlet)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    D = lambda x, y, t, u, p: 1 / 9
    return O.FVMProblem(mesh, BCs, diffusion_function=D, initial_condition=ic, final_time=0.5)


def example_heat_convection_problem(n=200):
    L, k, T0, Tinf, alpha, q, h = 1.0, 237.0, 10.0, 10.0, 80.0e-6, 10.0, 25.0
    tri = O.triangulate_rectangle(0, L, 0, L, n, n, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    bot = lambda x, y, t, T, p: -p["a"] * p["q"] / p["k"]
    right = lambda x, y, t, T, p: 0.0 * T
    top = lambda x, y, t, T, p: -p["a"] * p["h"] / p["k"] * (p["Tinf"] - T)
    left = lambda x, y, t, T, p: 0.0 * T
    params = (dict(a=alpha, q=q, k=k), None, dict(a=alpha, h=h, k=k, Tinf=Tinf), None)
    BCs = O.BoundaryConditions(mesh, (bot, right, top, left), (O.Neumann,) * 4, parameters=params)
    flux = lambda x, y, t, a, b, g, p: (-p["a"] * a, -p["a"] * b)
    ic = np.full(tri.num_points, T0)
    return O.FVMProblem(mesh, BCs, flux_function=flux, flux_parameters=dict(a=alpha),
                        initial_condition=ic, final_time=2000.0)


# ---- independent control-volume integral, test_functions.jl:1-34 and :601-657 ----------
def get_control_volume(tri, i):
    """Polygon of the control volume of node i, with the triangle each piece lies in."""
    P = tri.points
    p = P[i]
    pieces = []  # (point_a, point_b, triangle (i,j,k))
    # find a starting neighbour: boundary nodes start at the right boundary node
    start = None
    for (u, v), (s, e) in tri.boundary_edge_map.items():
        if u == i:
            start = v
    if start is None:
        start = next(v for (u, v) in tri.adjacent if u == i and tri.adjacent[(u, v)] >= 0)
    j = start
    first = True
    while True:
        k = tri.adjacent.get((i, j), -1)
        if k < 0:
            break
        q, r = P[j], P[k]
        c = (p + q + r) / 3
        pieces.append(((p + q) / 2, c, (i, j, k)))
        pieces.append((c, (p + r) / 2, (i, j, k)))
        j = k
        if j == start:
            break
    return pieces


def _on_same_side(a, b, L):
    if np.isclose(a[1], 0.0) and np.isclose(b[1], 0.0):
        return 1
    if np.isclose(a[0], L) and np.isclose(b[0], L):
        return 2
    if np.isclose(a[1], L) and np.isclose(b[1], L):
        return 3
    if np.isclose(a[0], 0.0) and np.isclose(b[0], 0.0):
        return 4
    return 0


def get_dudt_val(prob, u, t, i, is_diff=True):
    if i in prob.conditions.dirichlet_nodes:
        return 0.0
    mesh = prob.mesh
    tri = mesh.triangulation
    P = tri.points
    p = P[i]
    integ = 0.0
    pieces = get_control_volume(tri, i)

    def tri_abg(T):
        _, t_idx = mesh.safe_get_triangle_props(T)
        Ts = tuple(int(v) for v in tri.triangles[t_idx])
        s = mesh.s[t_idx]
        a = s[0] * u[Ts[0]] + s[1] * u[Ts[1]] + s[2] * u[Ts[2]]
        b = s[3] * u[Ts[0]] + s[4] * u[Ts[1]] + s[5] * u[Ts[2]]
        g = s[6] * u[Ts[0]] + s[7] * u[Ts[1]] + s[8] * u[Ts[2]]
        return a, b, g

    for (pa, pb, T) in pieces:
        Lseg = np.linalg.norm(pa - pb)
        nx, ny = (pb[1] - pa[1]) / Lseg, (pa[0] - pb[0]) / Lseg
        a, b, g = tri_abg(T)
        if is_diff:
            integ += Lseg * ((-a / 9) * nx + (-b / 9) * ny)
        else:
            integ += Lseg * ((-80.0e-6 * a) * nx + (-80.0e-6 * b) * ny)
    if not is_diff:
        # boundary pieces p -> midpoint on either boundary edge at i (test_functions.jl:627-643)
        for (uu, vv) in tri.boundary_edge_map:
            if uu != i and vv != i:
                continue
            other = vv if uu == i else uu
            m = (p + P[other]) / 2
            k = tri.adjacent[(uu, vv)]
            a, b, g = tri_abg((uu, vv, k))
            mx, my = (p + m) / 2
            side = _on_same_side(p, m, 1.0)
            if side == 1:
                q = -80.0e-6 * 10.0 / 237.0
            elif side == 3:
                q = -80.0e-6 * 25.0 / 237.0 * (10.0 - (a * mx + b * my + g))
            else:
                q = 0.0
            integ += np.linalg.norm(p - m) * q
    S = prob.source_function(p[0], p[1], t, u[i], prob.source_parameters)
    return S - integ / mesh.cv_volumes[i]


def test_dudt_val_diffusion():
    """test/equations.jl:13-23 -> test_dudt_val (test_functions.jl:645-657)."""
    prob = example_diffusion_problem()
    u = prob.initial_condition
    ref = np.array([get_dudt_val(prob, u, 0.0, i) for i in range(len(u))])
    du = O.fvm_eqs(np.zeros_like(u), u, prob, 0.0)
    assert isapprox(ref, du)
    rng = np.random.default_rng(1)
    u = 50 * rng.random(len(u))
    ref = np.array([get_dudt_val(prob, u, 0.0, i) for i in range(len(u))])
    du = O.fvm_eqs(np.zeros_like(u), u, prob, 0.0)
    assert isapprox(ref, du, rtol=1e-12)


def test_dudt_val_convection_and_corner():
    """test/equations.jl:25-92: all-Neumann convection problem on 200x200, and the exact
    corner-node hand calculation with (alpha,beta,gamma) == (0,0,10)."""
    prob = example_heat_convection_problem(200)
    mesh = prob.mesh
    tri = mesh.triangulation
    u = prob.initial_condition
    t = 0.0
    # (1,2,201) is a stored triangle (0-based (0,1,200)), equations.jl:42
    T = (0, 1, 200)
    assert T in mesh.triangle_props
    t_idx = mesh.triangle_props[T]
    assert O.get_shape_function_coefficients(mesh, t_idx, T, u) == (0.0, 0.0, 10.0)
    x1, y1 = tri.points[1, 0], tri.points[200, 1]
    poly = np.array([(0.0, 0.0), (x1 / 2, 0.0), (x1 / 3, y1 / 3), (0.0, y1 / 2)])
    area = 0.5 * abs(np.dot(poly[:, 0], np.roll(poly[:, 1], -1)) - np.dot(poly[:, 1], np.roll(poly[:, 0], -1)))
    assert math.isclose(mesh.cv_volumes[0], area, rel_tol=1e-12)
    # hand flux balance at node 1 (equations.jl:57-80)
    c1, c2, c3, c4 = poly
    m1, m2, m3, m4 = (c1 + c2) / 2, (c2 + c3) / 2, (c3 + c4) / 2, (c4 + c1) / 2
    l1, l4 = np.linalg.norm(c2 - c1), np.linalg.norm(c1 - c4)
    n2 = np.array([(c3 - c2)[1], -(c3 - c2)[0]])
    n3 = np.array([(c4 - c3)[1], -(c4 - c3)[0]])
    a, b, g = 0.0, 0.0, 10.0
    f1 = prob.conditions.functions[0](m1[0], m1[1], t, a * m1[0] + b * m1[1] + g) * l1
    f2 = np.dot(prob.flux_function(m2[0], m2[1], 0.0, a, b, g, prob.flux_parameters), n2)
    f3 = np.dot(prob.flux_function(m3[0], m3[1], 0.0, a, b, g, prob.flux_parameters), n3)
    f4 = prob.conditions.functions[3](m4[0], m4[1], t, a * m4[0] + b * m4[1] + g) * l4
    fl = -(1 / area) * (f1 + f2 + f3 + f4)
    du = O.fvm_eqs(np.zeros_like(u), u, prob, t)
    assert math.isclose(fl, du[0], rel_tol=1e-9)
    assert math.isclose(fl, get_dudt_val(prob, u, t, 0, False), rel_tol=1e-9)
    # whole-field check on a non-trivial u (sampled nodes keep the CPU suite short)
    rng = np.random.default_rng(2)
    u2 = 10 + rng.random(len(u))
    du2 = O.fvm_eqs(np.zeros_like(u2), u2, prob, t)
    nodes = np.r_[0:400, 39600:40000, rng.integers(0, 40000, 600), np.arange(0, 40000, 200), np.arange(199, 40000, 200)]
    ref = np.array([get_dudt_val(prob, u2, t, int(i), False) for i in nodes])
    assert isapprox(ref, du2[nodes], rtol=1e-10)
    # vectorised oracle == loop oracle bitwise
    assert np.array_equal(du2, O.fvm_eqs_vec(np.zeros_like(u2), u2, prob, t))


def test_shape_function_identities():
    """test/test_functions.jl:385-393 and test/geometry.jl:10-76."""
    prob = example_diffusion_problem()
    mesh = prob.mesh
    P, Tr = mesh.triangulation.points, mesh.triangulation.triangles
    rng = np.random.default_rng(3)
    u = rng.random(len(P))
    for t, (i, j, k) in enumerate(Tr.tolist()):
        s = mesh.s[t]
        M = np.array([[P[i, 0], P[i, 1], 1], [P[j, 0], P[j, 1], 1], [P[k, 0], P[k, 1], 1]])
        abg = np.linalg.solve(M, u[[i, j, k]])
        a, b, g = O.get_shape_function_coefficients(mesh, t, (i, j, k), u)
        assert np.allclose([a, b, g], abg, atol=1e-9)
        c = (P[i] + P[j] + P[k]) / 3
        assert math.isclose(s[0] * c[0] + s[3] * c[1] + s[6] + s[1] * c[0] + s[4] * c[1] + s[7]
                            + s[2] * c[0] + s[5] * c[1] + s[8], 1.0, rel_tol=1e-12)
        assert mesh.delta[t] > 0
    assert math.isclose(mesh.cv_volumes.sum(), 4.0, rel_tol=1e-13)
    h = 2.0 / 24
    assert math.isclose(mesh.cv_volumes[0], h * h / 6, rel_tol=1e-12)
    assert math.isclose(mesh.cv_volumes[24], h * h / 3, rel_tol=1e-12)
    assert math.isclose(mesh.cv_volumes[26], h * h, rel_tol=1e-12)


def test_cv_edge_geometry_direct():
    """test/test_functions.jl:404-436: cv-edge midpoints/normals/lengths vs direct formulas."""
    prob = example_heat_convection_problem(12)
    mesh = prob.mesh
    P, Tr = mesh.triangulation.points, mesh.triangulation.triangles
    for t, T in enumerate(Tr.tolist()):
        p, q, r = P[T[0]], P[T[1]], P[T[2]]
        c = (p + q + r) / 3
        for e, (i, j) in enumerate(((T[0], T[1]), (T[1], T[2]), (T[2], T[0]))):
            m = (P[i] + P[j]) / 2
            x, y = (c + m) / 2
            ex, ey = c - m
            l = math.hypot(ex, ey)
            got = mesh.get_cv_components(t, e)
            assert np.allclose(got, (x, y, ey / l, -ex / l, l), rtol=1e-13, atol=1e-15)


def test_single_triangle_sign_pattern():
    """test/test_functions.jl:527-556."""
    prob = example_diffusion_problem()
    u = 50 * np.random.default_rng(4).random(len(prob.initial_condition))
    mesh = prob.mesh
    for t_idx in range(0, len(mesh.triangulation.triangles), 7):
        T = tuple(int(v) for v in mesh.triangulation.triangles[t_idx])
        a, b, g = O.get_shape_function_coefficients(mesh, t_idx, T, u)
        q1, q2, q3 = (O.get_flux(prob, t_idx, a, b, g, 0.0, e) for e in range(3))
        du = np.zeros_like(u)
        O.fvm_eqs_single_triangle(du, u, prob, 0.0, t_idx)
        assert math.isclose(du[T[0]], -(q1 - q3), abs_tol=1e-9)
        assert math.isclose(du[T[1]], -(-q1 + q2), abs_tol=1e-9)
        assert math.isclose(du[T[2]], -(-q2 + q3), abs_tol=1e-9)


def test_eval_flux_function_exact():
    """test/problem.jl:30-35."""
    tri = O.triangulate_rectangle(0, 1, 0, 1, 3, 3, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0, O.Dirichlet)

    def flux(x, y, t, a, b, g, p):
        u = a * x + b * y + g
        return (-a * u * p[0] + t, x + t - b * u * p[1])

    prob = O.FVMProblem(mesh, BCs, flux_function=flux, flux_parameters=(-0.5, 1.3),
                        initial_condition=np.zeros(9), final_time=1.0)
    x, y, t, a, b, g = 0.5, -1.0, 2.3, 0.371, -5.37, 17.5
    u = a * x + b * y + g
    qx, qy = prob.eval_flux_function(x, y, t, a, b, g)
    assert qx == -a * u * (-0.5) + t
    assert qy == x + t - b * u * 1.3
    # construct_flux_function identity q = -D (alpha, beta), test/problem.jl:152-160
    D = lambda x, y, t, u, p: x + y + t + u + p
    q = O.construct_flux_function(None, D, 0.7)
    qx, qy = q(x, y, t, a, b, g, None)
    Dv = D(x, y, t, u, 0.7)
    assert (qx, qy) == (-Dv * a, -Dv * b)


def test_conditions_merge():
    """test/conditions.jl: section -> edge/node dictionaries, fidx offset nif, precedence."""
    tri = O.triangulate_rectangle(0, 1, 0, 1, 5, 4, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    f = lambda x, y, t, u, p: 0.0
    BCs = O.BoundaryConditions(mesh, (f, f, f, f), (O.Neumann, O.Dirichlet, O.Dudt, O.Constrained))
    ICs = O.InternalConditions((f,), dirichlet_nodes={6: 0}, dudt_nodes={7: 0})
    c = O.Conditions(mesh, BCs, ICs)
    assert c.neumann_edges == {(i, i + 1): 1 for i in range(4)}
    assert set(c.dirichlet_nodes) == {6, 4, 9, 14, 19} and c.dirichlet_nodes[4] == 2 and c.dirichlet_nodes[6] == 0
    assert c.dudt_nodes[7] == 0 and all(c.dudt_nodes[n] == 3 for n in (19, 18, 17, 16, 15))
    assert c.constrained_edges == {(15, 10): 4, (10, 5): 4, (5, 0): 4}
    assert len(c.functions) == 5
    # node 19 is both Dirichlet (right) and Dudt (top): Dirichlet wins in the node pass
    assert c.is_dirichlet_node(19) and c.is_dudt_node(19)


def test_readme_interior_row_and_template_consistency():
    """SURVEY 8c (ix),(x): interior row = D/h^2 [1,1,-4,1,1]; A u + b == fvm_eqs(u) away from
    conditioned nodes; MET b = -1 / identity rows."""
    tri = O.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)  # templates pass t=u=nothing
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    tpl = O.DiffusionEquation(mesh, BCs, diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.5)
    h = 2 / 49
    row = tpl.A[51 + 50].toarray().ravel()
    i = 101
    assert math.isclose(row[i], -4 / 9 / h**2, rel_tol=1e-12)
    for nb in (i - 1, i + 1, i - 50, i + 50):
        assert math.isclose(row[nb], 1 / 9 / h**2, rel_tol=1e-12)
    assert abs(row[i + 49]) < 1e-10 and abs(row[i - 49]) < 1e-10
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: 1 / 9, initial_condition=ic, final_time=0.5)
    u = 50 * np.random.default_rng(5).random(len(ic))
    du = O.fvm_eqs(np.zeros_like(u), u, prob, 0.0)
    assert isapprox(tpl.A @ u + tpl.b, du, rtol=1e-12)
    assert np.all(tpl.u0[list(tpl.conditions.dirichlet_nodes)] == 0.0)
    met = O.MeanExitTimeProblem(mesh, BCs, diffusion_function=lambda x, y, p: 1 / 9)
    dn = np.array(sorted(met.conditions.dirichlet_nodes))
    assert np.all(met.b[dn] == 0) and np.all(np.delete(met.b, dn) == -1)
    assert np.all(met.A.diagonal()[dn] == 1.0)


def test_poisson_closed_form():
    """docs/src/literate_wyos/poissons_equation.jl:92-116,159-162: rtol 1e-4 on 100x100."""
    tri = O.triangulate_rectangle(0, 1, 0, 1, 100, 100, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet)
    src = lambda x, y, p: -math.sin(math.pi * x) * math.sin(math.pi * y)
    tpl = O.PoissonsEquation(mesh, BCs, source_function=src)
    sol = O.solve_steady(tpl)
    P = tri.points
    exact = 1 / (2 * math.pi**2) * np.sin(math.pi * P[:, 0]) * np.sin(math.pi * P[:, 1])
    assert isapprox(sol, exact, rtol=1e-4)


def test_laplace_closed_form():
    """docs/src/literate_wyos/laplaces_equation.jl:160-193: u = 5 log6(1+x), rtol 1e-3."""
    tri = O.triangulate_rectangle(0, 5, 0, 5, 100, 100, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    zero_f = lambda x, y, t, u, p: 0.0
    five_f = lambda x, y, t, u, p: 5.0
    BCs = O.BoundaryConditions(mesh, (zero_f, five_f, zero_f, zero_f), (O.Neumann, O.Dirichlet, O.Neumann, O.Dirichlet))
    tpl = O.LaplacesEquation(mesh, BCs, diffusion_function=lambda x, y, p: (x + 1) * (y + 2))
    sol = O.solve_steady(tpl)
    exact = 5 * np.log(1 + tri.points[:, 0]) / math.log(6)
    assert isapprox(sol, exact, rtol=1e-3)


def test_tsit5_tableau_order_conditions_and_convergence():
    """row sums = c, order conditions 1..4, 5th-order convergence."""
    A, C = O.TSIT5_A, O.TSIT5_C
    for s in range(1, 7):
        assert math.isclose(sum(A[s]), C[s], abs_tol=1e-14)
    b = np.array(A[6] + (0.0,))
    c = np.array(C)
    Am = np.zeros((7, 7))
    for s in range(1, 7):
        Am[s, :s] = A[s]
    assert math.isclose(b.sum(), 1, abs_tol=1e-14)
    assert math.isclose(b @ c, 1 / 2, abs_tol=1e-14)
    assert math.isclose(b @ c**2, 1 / 3, abs_tol=1e-14)
    assert math.isclose(b @ (Am @ c), 1 / 6, abs_tol=1e-14)
    assert math.isclose(b @ c**3, 1 / 4, abs_tol=1e-14)
    assert math.isclose(b @ (c * (Am @ c)), 1 / 8, abs_tol=1e-14)
    assert math.isclose(b @ (Am @ c**2), 1 / 12, abs_tol=1e-14)
    assert math.isclose(b @ (Am @ (Am @ c)), 1 / 24, abs_tol=1e-14)
    assert math.isclose(b @ c**4, 1 / 5, abs_tol=1e-14)

    def f(du, u, t):
        du[...] = -u + np.sin(t)

    exact = lambda t: 1.5 * math.exp(-t) + 0.5 * (math.sin(t) - math.cos(t))
    errs = []
    for n in (10, 20, 40):
        u = O.tsit5_fixed(f, np.array([1.0]), 0.0, 1.0, 1.0 / n)
        errs.append(abs(u[0] - exact(1.0)))
    assert 4.5 < math.log2(errs[0] / errs[1]) < 6.5 and 4.5 < math.log2(errs[1] / errs[2]) < 6.5


def test_tsit5_embedded_error_weights():
    """The embedded estimator of the adaptive stepper (f3): btilde = b - bhat must annihilate every elementary weight up
    to order 4 (so bhat is a 4th-order method and the estimate is O(h^5)) and must NOT annihilate the 5th-order one;
    the same constants sit in fvm_solvers.cu (TS_BT).  Published values: Tsitouras 2011 / OrdinaryDiffEq's
    Tsit5ConstantCache -- neither is under /root/reference, so these identities are the pin that exists."""
    A, c = O.TSIT5_A, np.array(O.TSIT5_C)
    Am = np.zeros((7, 7))
    for s in range(1, 7):
        Am[s, :s] = A[s]
    bt = np.array(O.TSIT5_BTILDE)
    weights = [np.ones(7), c, c**2, Am @ c, c**3, c * (Am @ c), Am @ c**2, Am @ (Am @ c)]
    for w in weights:
        assert abs(bt @ w) <= 2e-15
    assert 1e-4 < abs(bt @ c**4) < 1e-3
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finitevolumemethod.jl_b200", "csrc", "fvm_solvers.cu")).read()
    m = re.search(r"TS_BT\[7\]\s*=\s*\{([^}]*)\}", src)
    assert m and np.array_equal(np.array([float(x) for x in m.group(1).replace("\n", " ").split(",")]), bt)


def test_system_equals_scalar():
    """test/equations.jl:106-121: a 2-species system of identical diffusion problems
    reproduces the scalar RHS in both species."""
    prob = example_diffusion_problem()
    mesh = prob.mesh
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0, O.Dirichlet)
    ic = prob.initial_condition
    q1 = lambda x, y, t, a, b, g, p: (-a[0] / 9, -b[0] / 9)
    q2 = lambda x, y, t, a, b, g, p: (-a[1] / 9, -b[1] / 9)
    p1 = O.FVMProblem(mesh, BCs, flux_function=q1, initial_condition=ic, final_time=0.5)
    p2 = O.FVMProblem(mesh, BCs, flux_function=q2, initial_condition=ic, final_time=0.5)
    sys_ = O.FVMSystem(p1, p2)
    rng = np.random.default_rng(6)
    u = 50 * rng.random(len(ic))
    U = np.stack([u, u], axis=1)
    dU = O.fvm_eqs(np.zeros_like(U), U, sys_, 0.0)
    du = O.fvm_eqs(np.zeros_like(u), u, prob, 0.0)
    assert isapprox(dU[:, 0], du, rtol=1e-13) and isapprox(dU[:, 1], du, rtol=1e-13)
    assert np.array_equal(dU, O.fvm_eqs_vec(np.zeros_like(U), U, sys_, 0.0))


def test_jacobian_sparsity_counts():
    """solve.jl:56-77: nnz = N + 2E with E = N + T - 1 on a lattice."""
    tri = O.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True)
    r, c = O.jacobian_sparsity(tri)
    assert len(r) == 17102


def test_c_oracle_matches_numpy_oracle():
    """oracle/fvm_oracle_c.c (the timed CPU arm) against the NumPy restatement: serial bitwise,
    threaded/flat to summation order (test/test_functions.jl:653-656 uses the same bar)."""
    from oracle.c_oracle import COracle
    prob = example_diffusion_problem()
    tri = prob.mesh.triangulation
    dn = np.array(sorted(prob.conditions.dirichlet_nodes), dtype=np.int32)
    co = COracle(tri.points, tri.triangles, dn, 1 / 9, nthreads=4)
    assert np.array_equal(co.volumes(), prob.mesh.cv_volumes)
    u = 50 * np.random.default_rng(8).random(tri.num_points)
    ref = O.fvm_eqs(np.zeros_like(u), u, prob, 0.0)
    assert np.array_equal(co.fvm_eqs_serial(u), ref)
    assert isapprox(co.fvm_eqs_threaded(u), ref, rtol=1e-14)
    assert isapprox(co.fvm_eqs_flat(u), ref, rtol=1e-14)


def test_general_c_oracle_matches_numpy_oracle():
    """oracle_fvm_eqs_general (the checker of the full-size GPU parity tests: registered flux / source forms,
    scalar problems and FVMSystems, geometry recomputed per triangle) is bitwise the NumPy oracle's serial
    fvm_eqs! wherever the boundary-edge pass adds nothing (all-Dirichlet boundary, homogeneous Neumann)."""
    import fvm_b200 as G
    from oracle import c_oracle
    from tests.common import Pair
    pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 23, 17, single_boundary=True))
    tri = pair.gtri
    N = tri.num_points
    rng = np.random.default_rng(3)
    u = 1 + rng.random(N)
    bnd = np.zeros(N, bool)
    bnd[np.unique(tri.boundary_edges()[0])] = True
    scalar = ((G.ConstantDiffusion(1 / 9), None, 0, [1 / 9], 0, []),
              (G.PowerDiffusion(0.3, 3.0), G.LogisticSource(1.3), 2, [0.3, 3.0, 0.0], 2, [1.3]),
              (G.PowerDiffusion(0.3, 1.0), G.LinearSource(-0.4, 0.2), 2, [0.3, 1.0, 0.0], 1, [-0.4, 0.2]),
              (G.AdvectionDiffusionFlux(0.05, 1.5, -0.7), None, 3, [0.05, 1.5, -0.7], 0, []))
    for flux, src, fm, fp, sm, sp in scalar:
        for bc, dirn in ((G.Dirichlet, bnd), (G.Neumann, None)):
            if bc is G.Neumann and fm == 3:
                continue  # advection through the boundary: the boundary-edge pass is not a no-op there
            gp, op = pair.problem(G.Const(0.0), bc, flux, source=src, ic=u)
            ref = O.fvm_eqs(np.zeros_like(u), u, op, 0.0)
            out = c_oracle.fvm_eqs_general(tri.points, tri.triangles, u, 1, fm, fp, sm, sp, dirichlet=dirn)
            assert np.array_equal(out, ref), (type(flux).__name__, bc)
    # 2-species systems: Keller-Segel (BASELINE config 4) and Gray-Scott, zero-flux Neumann
    U = np.ascontiguousarray(np.stack([0.5 + 0.5 * rng.random(N), 0.25 * rng.random(N)], axis=1))
    for fluxes, src, fm, fp, sm, sp in (((G.KellerSegelFlux(4.0, 1.0),) * 2, G.KellerSegelSource(0.1), 4, [4.0, 1.0], 6, [0.1]),
                                        ((G.ConstantDiffusion(2e-5), G.ConstantDiffusion(1e-5)), G.GrayScottSource(0.04, 0.1), 0,
                                         [2e-5, 1e-5], 4, [0.04, 0.1])):
        g1, o1 = pair.problem(G.Const(0.0), G.Neumann, fluxes[0], source=src, var=0, ic=U[:, 0])
        g2, o2 = pair.problem(G.Const(0.0), G.Neumann, fluxes[1], source=src, var=1, ic=U[:, 1])
        ref = O.fvm_eqs(np.zeros_like(U), U, O.FVMSystem(o1, o2), 0.0)
        out = c_oracle.fvm_eqs_general(tri.points, tri.triangles, U, 2, fm, fp, sm, sp)
        assert np.array_equal(out, ref)


def test_vectorised_template_assembly_matches_the_loop():
    """The vectorised MeanExitTimeProblem assembly (used only for the 10^6-unknown config-3 parity case) against the
    loop restatement of abstract_templates.jl:73-99 / mean_exit_time.jl:57-94: same pattern, values to summation order."""
    tri = O.triangulate_rectangle(0, 2, 0, 3, 31, 27, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, (lambda x, y, t, u, p: 0.0,) * 4, (O.Dirichlet, O.Neumann, O.Dirichlet, O.Neumann))
    Dfn = lambda x, y, p: 0.3 + 0.1 * x * y
    loop = O.MeanExitTimeProblem(mesh, BCs, diffusion_function=Dfn)
    vec = O.MeanExitTimeProblem(mesh, BCs, diffusion_function=Dfn, vectorised=True)
    assert np.array_equal(loop.b, vec.b)
    assert np.array_equal(loop.A.indptr, vec.A.indptr) and np.array_equal(loop.A.indices, vec.A.indices)
    assert np.abs(loop.A.data - vec.A.data).max() <= 4e-16 * np.abs(loop.A.data).max()


def test_readme_square_plate_series_solution():
    """BASELINE configs[0] end to end on the oracle: README diffusion (50x50, D = 1/9, Dirichlet 0) integrated
    with the fixed-step Tsit5 and the Dirichlet callback, against the separated-variables series of
    docs/src/literate_tutorials/diffusion_equation_on_a_square_plate.jl:79-94 (discretisation-level agreement;
    the tutorial only plots it)."""
    tri = O.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0 * u, O.Dirichlet)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: 1 / 9 + 0 * u, initial_condition=ic, final_time=0.5)
    u = O.tsit5_fixed(lambda d, x, t: O.fvm_eqs_vec(d, x, prob, t), ic, 0.0, 0.5, 0.0025,
                      callback=lambda x, t: (O.update_dirichlet_nodes(x, t, prob), True)[1])
    x, y = tri.points[:, 0], tri.points[:, 1]
    s = np.zeros_like(x)
    for m in range(1, 50, 2):
        mterm = 2 / m * np.sin(m * np.pi * x / 2) * np.exp(-np.pi**2 * m**2 * 0.5 / 36)
        for n in range(1, 50):
            s += mterm * (1 - np.cos(n * np.pi / 2)) / n * np.sin(n * np.pi * y / 2) * np.exp(-np.pi**2 * n**2 * 0.5 / 36)
    exact = 200 * s / np.pi**2
    assert np.abs(u - exact).max() <= 0.02 * exact.max()  # O(h^2) discretisation error on a 50x50 mesh
    assert 20.0 < exact.max() < 50.0


def test_brusselator_exact_solution_pins_system_path():
    """docs/src/literate_tutorials/reaction_diffusion_brusselator_system_of_pdes.jl:95-140,125-126: exact
    solution Phi = exp(-x-y-t/2), Psi = exp(x+y+t/2) of a 2-species FVMSystem with mixed time-dependent
    Neumann / Dirichlet data.  Pins the oracle's system RHS, Neumann edges, Dirichlet callback and sources."""
    tri = O.triangulate_rectangle(0, 1, 0, 1, 21, 21, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    P = tri.points
    e = math.e
    phi_bc = (lambda x, y, t, u, p: -1 / 4 * np.exp(-x - t / 2), lambda x, y, t, u, p: 1 / 4 * np.exp(-1 - y - t / 2),
              lambda x, y, t, u, p: np.exp(-1 - x - t / 2), lambda x, y, t, u, p: -1 / 4 * np.exp(-y - t / 2))
    psi_bc = (lambda x, y, t, u, p: np.exp(x + t / 2), lambda x, y, t, u, p: -1 / 4 * np.exp(1 + y + t / 2),
              lambda x, y, t, u, p: -1 / 4 * np.exp(1 + x + t / 2), lambda x, y, t, u, p: np.exp(y + t / 2))
    pb = O.BoundaryConditions(mesh, phi_bc, (O.Neumann, O.Neumann, O.Dirichlet, O.Neumann))
    sb = O.BoundaryConditions(mesh, psi_bc, (O.Dirichlet, O.Neumann, O.Neumann, O.Dirichlet))
    p1 = O.FVMProblem(mesh, pb, flux_function=lambda x, y, t, a, b, g, p: (-a[0] / 4, -b[0] / 4),
                      source_function=lambda x, y, t, u, p: u[0] ** 2 * u[1] - 2 * u[0], initial_condition=np.exp(-P[:, 0] - P[:, 1]),
                      final_time=0.5)
    p2 = O.FVMProblem(mesh, sb, flux_function=lambda x, y, t, a, b, g, p: (-a[1] / 4, -b[1] / 4),
                      source_function=lambda x, y, t, u, p: -u[0] ** 2 * u[1] + u[0], initial_condition=np.exp(P[:, 0] + P[:, 1]),
                      final_time=0.5)
    sys_ = O.FVMSystem(p1, p2)
    u = O.tsit5_fixed(lambda d, x, t: O.fvm_eqs_vec(d, x, sys_, t), sys_.initial_condition, 0.0, 0.5, 2e-3,
                      callback=lambda x, t: (O.update_dirichlet_nodes(x, t, sys_), True)[1])
    exact = np.stack([np.exp(-P[:, 0] - P[:, 1] - 0.25), np.exp(P[:, 0] + P[:, 1] + 0.25)], axis=1)
    assert np.abs(u - exact).max() <= 2e-3 * np.abs(exact).max()


def test_helmholtz_inhomogeneous_neumann_closed_form():
    """docs/src/literate_tutorials/helmholtz_equation_with_inhomogeneous_boundary_conditions.jl:20-45,82-84: the
    steady state of u_t = lap(u) + u with q.n = -1 on [-1,1]^2 is u = -(cos(x+1)+cos(1-x)+cos(y+1)+cos(1-y))/sin(2).
    The RHS is affine in u, so one Newton step from a Jacobian assembled column by column from the ORACLE's
    fvm_eqs! is exact.  Pins the inhomogeneous Neumann edges and the u-dependent source (the tutorial only
    compares images)."""
    n = 25
    tri = O.triangulate_rectangle(-1, 1, -1, 1, n, n, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: -1.0 + 0.0 * u, O.Neumann)
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: 1.0 + 0.0 * u, source_function=lambda x, y, t, u, p: u,
                        initial_condition=np.zeros(n * n), final_time=np.inf)
    N = n * n
    f0 = O.fvm_eqs_vec(np.zeros(N), np.zeros(N), prob, 0.0).copy()
    J = np.empty((N, N))
    for j in range(N):
        e = np.zeros(N)
        e[j] = 1.0
        J[:, j] = O.fvm_eqs_vec(np.zeros(N), e, prob, 0.0) - f0
    u = np.linalg.solve(J, -f0)
    assert np.abs(O.fvm_eqs_vec(np.zeros(N), u, prob, 0.0)).max() <= 1e-9 * np.abs(f0).max()
    x, y = tri.points[:, 0], tri.points[:, 1]
    exact = -(np.cos(x + 1) + np.cos(1 - x) + np.cos(y + 1) + np.cos(1 - y)) / math.sin(2)
    assert np.abs(u - exact).max() <= 2e-3 * np.abs(exact).max()  # O(h^2) at h = 1/12: measured 1.03e-3
    assert -2.6 < exact.min() < exact.max() < -0.9  # the level range the tutorial plots (-2.5:0.15:-1.0)


def test_porous_medium_barenblatt_solution():
    """docs/src/literate_tutorials/porous_medium_equation.jl:20-50,88-99: u_t = div(D u^(m-1) grad u) has the
    Barenblatt similarity solution.  Started from the exact profile at t = 1 (instead of the tutorial's narrow
    Gaussian, which needs an implicit integrator) and integrated to t = 2 with the fixed-step Tsit5: pins the
    u-dependent diffusion path D(u) = D0 u^(m-1) of construct_flux_function (problem.jl:425-440)."""
    m, M, D = 2, 0.37, 2.53
    RmM = 4 * m / (m - 1) * (M / (4 * np.pi)) ** ((m - 1) / m)

    def exact(x, y, t):
        r2 = x * x + y * y
        inner = (M / (4 * np.pi)) ** ((m - 1) / m) - (m - 1) / (4 * m) * r2 * (D * t) ** (-1 / m)
        return np.where(r2 < RmM * (D * t) ** (1 / m), (D * t) ** (-1 / m) * np.maximum(inner, 0.0) ** (1 / (m - 1)), 0.0)

    L = 3.0
    tri = O.triangulate_rectangle(-L, L, -L, L, 61, 61, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    x, y = tri.points[:, 0], tri.points[:, 1]
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: 0.0 * u, O.Dirichlet)
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: p[0] * u ** (p[1] - 1), diffusion_parameters=(D, m),
                        initial_condition=exact(x, y, 1.0), initial_time=1.0, final_time=2.0)
    u = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, prob, t), prob.initial_condition, 1.0, 2.0, 0.005,
                      callback=lambda v, t: (O.update_dirichlet_nodes(v, t, prob), True)[1])
    ref = exact(x, y, 2.0)
    assert np.abs(u - ref).max() <= 0.03 * ref.max()  # the free boundary is resolved to O(h)
    assert abs(u @ mesh.cv_volumes - M) <= 2e-3 * M  # mass M is conserved (zero flux through the support)
    assert (ref > 0).sum() > 500 and ref[0] == 0.0


def _wedge_series(r, t, terms=40):
    """Exact solution of the wedge tutorial for f = 1 - r (diffusion_equation_in_a_wedge_with_mixed_boundary_conditions.jl
    :112-150): f does not depend on theta, so only the order-0 terms of the tutorial's double series survive:
    u = sum_m [2 I_m / J1(z_m)^2] exp(-z_m^2 t) J0(z_m r),  I_m = int_0^1 (1-s) s J0(z_m s) ds,  J0(z_m) = 0."""
    from scipy.integrate import quad
    from scipy.special import j0, j1, jn_zeros
    z = jn_zeros(0, terms)
    c = [2 * quad(lambda s: (1 - s) * s * j0(zm * s), 0, 1)[0] / j1(zm) ** 2 for zm in z]
    return sum(cm * np.exp(-zm * zm * t) * j0(zm * r) for cm, zm in zip(c, z))


def test_wedge_mixed_conditions_bessel_series():
    """docs/src/literate_tutorials/diffusion_equation_in_a_wedge_with_mixed_boundary_conditions.jl:20-45 (alpha = pi/4,
    zero-flux Neumann on the two straight edges, Dirichlet u = 0 on the arc, f = 1 - r, D = 1, t = 0.1) against the
    tutorial's exact series (:133-150): pins three boundary sections of different kinds on a curved domain, the
    homogeneous Neumann edges and the Dirichlet callback."""
    from tests.common import to_oracle_tri, wedge_mesh
    tri = to_oracle_tri(wedge_mesh(24))
    mesh = O.FVMGeometry(tri)
    r = np.hypot(tri.points[:, 0], tri.points[:, 1])
    zero = lambda x, y, t, u, p: 0.0 * x
    BCs = O.BoundaryConditions(mesh, (zero, zero, zero), (O.Neumann, O.Dirichlet, O.Neumann))
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: 1.0 + 0.0 * u, initial_condition=1 - r, final_time=0.1)
    u = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, prob, t), prob.initial_condition, 0.0, 0.1, 2e-4,
                      callback=lambda v, t: (O.update_dirichlet_nodes(v, t, prob), True)[1])
    exact = _wedge_series(r, 0.1)
    assert np.abs(u - exact).max() <= 4e-3 * exact.max()  # measured 1.9e-3 at 24 rings (1.1e-3 at 32)
    assert np.abs(u[r > 1 - 1e-12]).max() == 0.0  # the arc stays at its Dirichlet value
    # the template assembles the same semi-discrete operator (diffusion_equation.jl:69-101): A u + b == fvm_eqs!(u)
    tpl = O.DiffusionEquation(mesh, BCs, diffusion_function=lambda x, y, p: 1.0, initial_condition=1 - r, final_time=0.1)
    free = r < 1 - 1e-12
    du = O.fvm_eqs_vec(np.zeros_like(u), u, prob, 0.05)
    assert np.abs((tpl.A @ u + tpl.b - du)[free]).max() <= 1e-11 * np.abs(du).max()


def test_disk_dudt_boundary_exact_solution():
    """docs/src/literate_tutorials/reaction_diffusion_equation_with_a_time_dependent_dirichlet_boundary_condition_on_a_disk.jl
    :20-45: u_t = div(u grad u) + u(1-u) on the unit disk with du/dt = u on the boundary has the exact solution
    u = exp(t) sqrt(I0(sqrt(2) r)) (:62-65 of the tutorial's test block).  Pins Dudt boundary nodes (source_contributions.jl
    :10-12), the u-dependent diffusion and the nonlinear source on an unstructured mesh."""
    from scipy.special import i0
    from tests.common import disk_mesh, to_oracle_tri
    tri = to_oracle_tri(disk_mesh(16))
    mesh = O.FVMGeometry(tri)
    r = np.hypot(tri.points[:, 0], tri.points[:, 1])
    ic = np.sqrt(i0(np.sqrt(2) * r))
    BCs = O.BoundaryConditions(mesh, lambda x, y, t, u, p: u, O.Dudt)
    prob = O.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: u, source_function=lambda x, y, t, u, p: u * (1 - u),
                        initial_condition=ic, final_time=0.1)
    u = O.tsit5_fixed(lambda d, v, t: O.fvm_eqs_vec(d, v, prob, t), ic, 0.0, 0.1, 5e-4)
    exact = math.exp(0.1) * ic
    assert np.abs(u - exact).max() <= 2e-4 * exact.max()  # measured 5.7e-5
    bnd = r > 1 - 1e-9
    assert bnd.sum() == 96 and np.abs(u[bnd] - exact[bnd]).max() <= 1e-12 * exact.max()  # du/dt = u integrated to Tsit5 accuracy


def test_control_volume_polygon_areas_unstructured():
    """test/geometry.jl:16-20: `geo.cv_volumes[i]` equals the area of the control-volume polygon around vertex i (triangle
    centroids joined to edge midpoints).  Restated independently of geometry.jl:117-130 (which uses cross products of
    centroid-to-vertex and midpoint-to-midpoint vectors): shoelace area of the quadrilateral vertex -> midpoint of the next
    edge -> centroid -> midpoint of the previous edge in every incident triangle, on an unstructured Delaunay mesh with
    points that are not vertices (volume 0, fix_missing_vertices)."""
    from tests.common import delaunay_mesh, to_oracle_tri
    tri = to_oracle_tri(delaunay_mesh(400, 11, extra_points=3))
    mesh = O.FVMGeometry(tri)
    P, Tr = tri.points, tri.triangles

    def shoelace(poly):
        x, y = np.asarray(poly).T
        return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))

    vol = np.zeros(len(P))
    for T in Tr.tolist():
        c = P[T].mean(axis=0)
        for r in range(3):
            v, nxt, prv = T[r], T[(r + 1) % 3], T[(r + 2) % 3]
            vol[v] += shoelace([P[v], (P[v] + P[nxt]) / 2, c, (P[v] + P[prv]) / 2])
    assert np.allclose(mesh.cv_volumes, vol, rtol=1e-10, atol=0.0)  # the shoelace sums cancel in absolute coordinates: 1.5e-12 measured
    assert np.all(mesh.cv_volumes[-3:] == 0.0) and np.all(mesh.cv_volumes[:-3] > 0.0)
    # the control volumes tile the domain: their total is the total triangle area
    p, q, r = P[Tr[:, 0]], P[Tr[:, 1]], P[Tr[:, 2]]
    area = 0.5 * ((q[:, 0] - p[:, 0]) * (r[:, 1] - p[:, 1]) - (q[:, 1] - p[:, 1]) * (r[:, 0] - p[:, 0]))
    assert math.isclose(mesh.cv_volumes.sum(), area.sum(), rel_tol=1e-12)


def test_compute_flux_and_pl_interpolate_restated():
    """test/test_functions.jl:659-750 (test_compute_flux): on every boundary edge (i, j) and every interior edge,
    compute_flux(prob, i, j, u, t) (problem.jl:458-487) equals q(midpoint) . n with n = (e_y, -e_x)/|e| the normal to the
    RIGHT of i -> j, where alpha, beta, gamma come from an independent 3x3 solve on the triangle that holds the edge (the
    triangle left of i -> j for a boundary edge, left of j -> i for an interior one); the system form returns one value per
    species and its first species equals the scalar problem's value.  pl_interpolate (utils.jl:23-27) reproduces the nodal
    values at the vertices and is affine inside the triangle."""
    tri = O.triangulate_rectangle(0.0, 1.0, 0.0, 1.5, 6, 5, single_boundary=False)
    mesh = O.FVMGeometry(tri)
    zero = lambda x, y, t, u, p: 0.0 * x
    BCs = O.BoundaryConditions(mesh, (zero,) * 4, (O.Neumann, O.Dirichlet, O.Dudt, O.Neumann))
    P, Tr = tri.points, tri.triangles
    rng = np.random.default_rng(12)
    u = rng.random(len(P))
    qfun = lambda x, y, t, a, b, g, p: (-a * x + t * g, x + t - b * (a * x + b * y + g) * p[1])  # test_functions.jl:287-293 style
    prob = O.FVMProblem(mesh, BCs, flux_function=qfun, flux_parameters=(0.5, 1.3), initial_condition=u, final_time=5.0)
    q1 = lambda x, y, t, a, b, g, p: qfun(x, y, t, a[0], b[0], g[0], p)
    q2 = lambda x, y, t, a, b, g, p: (-a[1] * b[0], g[1] - x * b[1])
    sys_ = O.FVMSystem(O.FVMProblem(mesh, BCs, flux_function=q1, flux_parameters=(0.5, 1.3), initial_condition=u, final_time=5.0),
                       O.FVMProblem(mesh, BCs, flux_function=q2, initial_condition=u, final_time=5.0))
    U = np.stack([u, rng.random(len(P))], axis=1)
    left_of = {}  # directed edge -> third vertex of the triangle on its left
    for a, b, c in Tr.tolist():
        left_of[(a, b)], left_of[(b, c)], left_of[(c, a)] = c, a, b

    def abg(vals, verts):
        M = np.array([[P[v, 0], P[v, 1], 1.0] for v in verts])
        return np.linalg.solve(M, vals[list(verts)])

    n_bnd = n_int = 0
    for (i, j), k in left_of.items():
        boundary = (j, i) not in left_of
        if boundary:
            verts = (i, j, k)  # get_adjacent(tri, i, j)
            n_bnd += 1
        else:
            verts = (j, i, left_of[(j, i)])  # get_adjacent(tri, j, i): the triangle on the other side
            n_int += 1
        p, q = P[i], P[j]
        ex, ey = (q - p) / np.linalg.norm(q - p)
        nx, ny = ey, -ex
        assert (q[0] - p[0]) * (ny) - (q[1] - p[1]) * (nx) < 0  # midpoint + n lies to the right of p -> q
        mx, my = (p + q) / 2
        a, b, g = abg(u, verts)
        qv = qfun(mx, my, 2.5, a, b, g, (0.5, 1.3))
        got = O.compute_flux(prob, i, j, u, 2.5)
        assert math.isclose(got, qv[0] * nx + qv[1] * ny, rel_tol=1e-10, abs_tol=1e-12)
        A = [abg(U[:, v], verts) for v in range(2)]
        al, be, ga = [A[0][0], A[1][0]], [A[0][1], A[1][1]], [A[0][2], A[1][2]]
        want = [np.dot(f(mx, my, 2.5, al, be, ga, (0.5, 1.3)), (nx, ny)) for f in (q1, q2)]
        gs = O.compute_flux(sys_, i, j, U, 2.5)
        assert np.allclose(gs, want, rtol=1e-10, atol=1e-12) and math.isclose(gs[0], got, rel_tol=1e-10, abs_tol=1e-12)
    assert n_bnd == 2 * (5 + 4) and n_int == 2 * (len(Tr) * 3 - n_bnd) // 2
    for T in Tr.tolist():
        for v in T:
            assert math.isclose(O.pl_interpolate(prob, T, u, P[v, 0], P[v, 1]), u[v], rel_tol=1e-11, abs_tol=1e-12)
        w = rng.dirichlet((1, 1, 1))
        x, y = w @ P[T]
        assert math.isclose(O.pl_interpolate(prob, T, u, x, y), w @ u[T], rel_tol=1e-10, abs_tol=1e-12)
        Ts = (T[1], T[2], T[0])  # any rotation of the stored key finds the same triangle (_safe_get_triangle_props, utils.jl:1-14)
        assert math.isclose(O.pl_interpolate(prob, Ts, u, x, y), w @ u[T], rel_tol=1e-10, abs_tol=1e-12)
        vals = O.pl_interpolate(sys_, T, U, x, y)
        assert np.allclose(vals, w @ U[T], rtol=1e-10, atol=1e-12)


def test_jacobian_sparsity_pattern_restated():
    """test/test_functions.jl:749-794 (test_jacobian_sparsity): the pattern is A[i,i] = 1 and A[i,j] = 1 for every
    non-ghost neighbour j of i (scalar); for an N-species system the unknowns are interleaved node-major
    (idx = (node-1) N + species, :768-779) and every (species, species) pair of a connected node pair is set.  The
    neighbour relation is restated here straight from the triangles, on an unstructured mesh."""
    from tests.common import delaunay_mesh, to_oracle_tri
    tri = to_oracle_tri(delaunay_mesh(150, 4))
    n = tri.num_points
    adj = np.zeros((n, n), dtype=bool)
    for a, b, c in tri.triangles.tolist():
        for i, j in ((a, b), (b, c), (c, a)):
            adj[i, j] = adj[j, i] = True
    np.fill_diagonal(adj, True)
    r, c = O.jacobian_sparsity(tri)
    got = np.zeros((n, n), dtype=bool)
    got[r, c] = True
    assert np.array_equal(got, adj) and len(r) == adj.sum()  # no duplicates
    for N in (2, 3):
        R, C = O.jacobian_sparsity(tri, N)
        want = np.kron(adj, np.ones((N, N), dtype=bool))  # node-major interleaving: block (i, j) is all ones
        got = np.zeros((n * N, n * N), dtype=bool)
        got[R, C] = True
        assert np.array_equal(got, want) and len(R) == want.sum()
