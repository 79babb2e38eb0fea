"""CPU ORACLE for the FiniteVolumeMethod.jl hot path.  TEST INFRASTRUCTURE ONLY.

This file is a line-by-line CPU restatement (NumPy + plain Python loops) of the
reference's semi-discrete right-hand side ``fvm_eqs!`` and of its linear
template assembly.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product (``libfvmcuda.so`` and the
``finitevolumemethod.jl_b200`` host package) never imports or calls anything
in ``oracle/``.

Parity status
-------------
Julia is not installed in the authoring container, so the reference itself
cannot be executed.  The oracle is pinned against every
numeric known-answer check the reference's own tests hold for this path
(``tests/test_oracle_golden.py``): the independent control-volume integral
``get_dudt_val`` (/root/reference/test/test_functions.jl:601-657), the
triangle sign pattern (:546-553), the corner-node hand calculation with
``(alpha,beta,gamma) == (0,0,10)`` exactly (/root/reference/test/equations.jl:37-92),
the shape-function identities (test_functions.jl:385-393), the exact
``eval_flux_function`` tuple (/root/reference/test/problem.jl:30-34) and the closed-form
Poisson / Laplace answers (docs/src/literate_wyos/poissons_equation.jl:113-116,
laplaces_equation.jl:188-193).  What stays **parity unpinned** because the
arithmetic lives in third-party Julia packages absent from /root/reference:
the lattice numbering of DelaunayTriangulation.triangulate_rectangle (only
corroborated by the tests cited in SURVEY.md Appendix B), the Tsit5 tableau of
OrdinaryDiffEq (checked here against the order conditions only) and the
callback/FSAL interplay.

Index convention: everything here is 0-based; the reference is 1-based.
Triangle ``(1,2,201)`` of the reference is ``(0,1,200)`` here.  A system's
state ``u`` has shape ``(N, neq)`` in C order, which is byte-identical to the
reference's column-major ``Matrix(neq, N)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Condition types  (/root/reference/src/conditions.jl:36-41)
# --------------------------------------------------------------------------
Neumann, Dudt, Dirichlet, Constrained = "Neumann", "Dudt", "Dirichlet", "Constrained"


# --------------------------------------------------------------------------
# Mesh.  Restates what the reference uses of DelaunayTriangulation.jl
# (third-party, compat 1.6.6, not under /root/reference).
# --------------------------------------------------------------------------
class Triangulation:
    """Minimal triangulation: points, ccw triangles (stored rotation matters),
    ccw boundary sections, and the adjacent map ``(u,v) -> w``.

    Ghost vertex of boundary section ``s`` (0-based) is ``-(s+1)``, i.e. the
    reference's ``-s`` for 1-based sections (test/test_functions.jl:198-199).
    """

    def __init__(self, points, triangles, boundary_sections=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int64).reshape(-1, 3)
        self.adjacent: Dict[Tuple[int, int], int] = {}
        for (i, j, k) in self.triangles.tolist():
            self.adjacent[(i, j)] = k
            self.adjacent[(j, k)] = i
            self.adjacent[(k, i)] = j
        if boundary_sections is None:
            boundary_sections = self._chain_boundary()
        self.boundary_sections = [np.asarray(s, dtype=np.int64) for s in boundary_sections]
        # boundary edge map: (u,v) -> (section, position); ghost on the (v,u) side
        self.boundary_edge_map: Dict[Tuple[int, int], Tuple[int, int]] = {}
        for s, nodes in enumerate(self.boundary_sections):
            for e in range(len(nodes) - 1):
                u, v = int(nodes[e]), int(nodes[e + 1])
                self.boundary_edge_map[(u, v)] = (s, e)
                self.adjacent[(v, u)] = -(s + 1)
        self._vertices = np.zeros(len(self.points), dtype=bool)
        self._vertices[self.triangles.ravel()] = True

    def _chain_boundary(self):
        """Boundary = directed triangle edges (u,v) with no (v,u) partner;
        chained into closed ccw loops, one section per loop."""
        nxt = {}
        for (u, v) in self.adjacent:
            if (v, u) not in self.adjacent:
                nxt[u] = v
        sections = []
        seen = set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop = [start]
            seen.add(start)
            cur = nxt[start]
            while cur != start:
                loop.append(cur)
                seen.add(cur)
                cur = nxt[cur]
            loop.append(start)
            sections.append(loop)
        return sections

    @property
    def num_points(self):
        return len(self.points)

    def has_vertex(self, i):
        return bool(self._vertices[i])

    def each_solid_vertex(self):
        return np.nonzero(self._vertices)[0]

    def get_adjacent(self, u, v):
        return self.adjacent[(u, v)]

    def get_neighbours(self):
        """node -> sorted set of solid neighbours (for jacobian_sparsity)."""
        nb = [set() for _ in range(self.num_points)]
        for (i, j, k) in self.triangles.tolist():
            nb[i].update((j, k))
            nb[j].update((i, k))
            nb[k].update((i, j))
        return nb


def triangulate_rectangle(a, b, c, d, nx, ny, single_boundary=False):
    """Lattice of SURVEY Appendix B (DelaunayTriangulation.triangulate_rectangle).

    Points ``idx(i,j) = i + j*nx`` (x fastest); per cell two ccw triangles
    ``(idx(i,j), idx(i+1,j), idx(i,j+1))`` and ``(idx(i,j+1), idx(i+1,j),
    idx(i+1,j+1))``; boundary ccw bottom,right,top,left.  Corroborated by
    /root/reference/test/equations.jl:42-56 (triangle (1,2,201) for nx=200) and
    test/test_functions.jl:112,350.
    """
    dx = (b - a) / (nx - 1)
    dy = (d - c) / (ny - 1)
    pts = np.empty((nx * ny, 2))
    ii = np.arange(nx, dtype=np.float64)
    jj = np.arange(ny, dtype=np.float64)
    pts[:, 0] = np.tile(a + ii * dx, ny)
    pts[:, 1] = np.repeat(c + jj * dy, nx)
    I, J = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    I = I.ravel()
    J = J.ravel()  # j outer, i inner
    p00 = I + J * nx
    p10 = p00 + 1
    p01 = p00 + nx
    p11 = p01 + 1
    tris = np.empty((2 * len(p00), 3), dtype=np.int64)
    tris[0::2] = np.stack([p00, p10, p01], axis=1)
    tris[1::2] = np.stack([p01, p10, p11], axis=1)
    bottom = np.arange(0, nx)
    right = np.arange(nx - 1, nx * ny, nx)
    top = np.arange(nx * ny - 1, nx * (ny - 1) - 1, -1)
    left = np.arange(nx * (ny - 1), -1, -nx)
    if single_boundary:
        sections = [np.concatenate([bottom, right[1:], top[1:], left[1:]])]
    else:
        sections = [bottom, right, top, left]
    return Triangulation(pts, tris, sections)


# --------------------------------------------------------------------------
# FVMGeometry  (/root/reference/src/geometry.jl:99-169)
# --------------------------------------------------------------------------
class FVMGeometry:
    """cv_volumes + per-triangle properties, arithmetic in the reference order.

    ``s`` (T,9) shape-function coefficients, ``mid`` (T,3,2) cv-edge midpoints,
    ``nrm`` (T,3,2) cv-edge unit normals, ``len`` (T,3) cv-edge lengths.
    ``triangle_props`` maps the stored vertex triple to its row (the reference's
    Dict, geometry.jl:44-49).
    """

    def __init__(self, tri: Triangulation):
        self.triangulation = tri
        P, Tr = tri.points, tri.triangles
        i, j, k = Tr[:, 0], Tr[:, 1], Tr[:, 2]
        px, py = P[i, 0], P[i, 1]
        qx, qy = P[j, 0], P[j, 1]
        rx, ry = P[k, 0], P[k, 1]
        # centroid and edge midpoints (geometry.jl:114-115; DelaunayTriangulation
        # statistics: (p+q+r)/3 and (p+q)/2)
        cx, cy = (px + qx + rx) / 3, (py + qy + ry) / 3
        m1x, m1y = (px + qx) / 2, (py + qy) / 2
        m2x, m2y = (qx + rx) / 2, (qy + ry) / 2
        m3x, m3y = (rx + px) / 2, (ry + py) / 2
        # sub-control-volume areas (geometry.jl:118-135)
        pcx, pcy = cx - px, cy - py
        qcx, qcy = cx - qx, cy - qy
        rcx, rcy = cx - rx, cy - ry
        m13x, m13y = m1x - m3x, m1y - m3y
        m21x, m21y = m2x - m1x, m2y - m1y
        m32x, m32y = m3x - m2x, m3y - m2y
        S1 = 1 / 2 * np.abs(pcx * m13y - pcy * m13x)
        S2 = 1 / 2 * np.abs(qcx * m21y - qcy * m21x)
        S3 = 1 / 2 * np.abs(rcx * m32y - rcy * m32x)
        V = np.zeros(tri.num_points)
        # sequential accumulation in triangle order, like the serial loop
        contrib = np.stack([S1, S2, S3], axis=1).ravel()
        np.add.at(V, Tr.ravel(), contrib)
        self.cv_volumes = V
        # shape function coefficients (geometry.jl:137-146), left-to-right
        D = qx * ry - qy * rx - px * ry + rx * py + px * qy - qx * py
        s = np.empty((len(Tr), 9))
        s[:, 0] = (qy - ry) / D
        s[:, 1] = (ry - py) / D
        s[:, 2] = (py - qy) / D
        s[:, 3] = (rx - qx) / D
        s[:, 4] = (px - rx) / D
        s[:, 5] = (qx - px) / D
        s[:, 6] = (qx * ry - rx * qy) / D
        s[:, 7] = (rx * py - px * ry) / D
        s[:, 8] = (px * qy - qx * py) / D
        self.s = s
        self.delta = D
        # cv-edge midpoints, normals, lengths (geometry.jl:149-161)
        mid = np.empty((len(Tr), 3, 2))
        nrm = np.empty((len(Tr), 3, 2))
        ln = np.empty((len(Tr), 3))
        for e, (mx, my) in enumerate(((m1x, m1y), (m2x, m2y), (m3x, m3y))):
            mid[:, e, 0] = (mx + cx) / 2
            mid[:, e, 1] = (my + cy) / 2
            ex, ey = cx - mx, cy - my
            l = np.sqrt(ex * ex + ey * ey)  # LinearAlgebra.norm of a 2-tuple, unscaled branch
            ln[:, e] = l
            nrm[:, e, 0] = ey / l
            nrm[:, e, 1] = -ex / l
        self.mid, self.nrm, self.len = mid, nrm, ln
        self.triangle_props: Dict[Tuple[int, int, int], int] = {
            (int(a), int(b), int(c)): t for t, (a, b, c) in enumerate(Tr.tolist())
        }

    # utils.jl:1-14
    def safe_get_triangle_props(self, T):
        i, j, k = T
        if (i, j, k) in self.triangle_props:
            return (i, j, k), self.triangle_props[(i, j, k)]
        elif (j, k, i) in self.triangle_props:
            return (j, k, i), self.triangle_props[(j, k, i)]
        else:
            return (k, i, j), self.triangle_props[(k, i, j)]

    def get_cv_components(self, t, e):  # control_volumes.jl:16-21
        return self.mid[t, e, 0], self.mid[t, e, 1], self.nrm[t, e, 0], self.nrm[t, e, 1], self.len[t, e]

    def get_boundary_cv_components(self, i, j):  # control_volumes.jl:41-56
        P = self.triangulation.points
        px, py = P[i]
        qx, qy = P[j]
        lij = math.sqrt((qx - px) * (qx - px) + (qy - py) * (qy - py))
        nx, ny = (qy - py) / lij, -(qx - px) / lij
        mijx, mijy = (px + qx) / 2, (py + qy) / 2
        mix, miy = (px + mijx) / 2, (py + mijy) / 2
        mjx, mjy = (qx + mijx) / 2, (qy + mijy) / 2
        li = math.sqrt((mijx - px) * (mijx - px) + (mijy - py) * (mijy - py))
        k = self.triangulation.get_adjacent(i, j)
        T, t = self.safe_get_triangle_props((i, j, k))
        return nx, ny, mix, miy, mjx, mjy, li, T, t


# --------------------------------------------------------------------------
# Conditions  (/root/reference/src/conditions.jl)
# --------------------------------------------------------------------------
def _wrap(fnc, p):  # ParametrisedFunction, conditions.jl:16-20
    return lambda x, y, t, u: fnc(x, y, t, u, p)


class BoundaryConditions:  # conditions.jl:164-172,237-251
    def __init__(self, mesh, functions, condition_types, parameters=None):
        if callable(functions):
            functions = (functions,)
        if isinstance(condition_types, str):
            condition_types = (condition_types,)
        nsec = len(mesh.triangulation.boundary_sections)
        assert len(functions) == len(condition_types) == nsec, \
            "The number of boundary condition functions must match the number of boundary sections."
        if parameters is None:
            parameters = (None,) * len(functions)
        self.functions = tuple(_wrap(f, p) for f, p in zip(functions, parameters))
        self.condition_types = tuple(condition_types)


class InternalConditions:  # conditions.jl:223-230,253-269
    def __init__(self, functions=(), dirichlet_nodes=None, dudt_nodes=None, parameters=None):
        if callable(functions):
            functions = (functions,)
        if parameters is None:
            parameters = (None,) * len(functions)
        self.functions = tuple(_wrap(f, p) for f, p in zip(functions, parameters))
        self.dirichlet_nodes = dict(dirichlet_nodes or {})
        self.dudt_nodes = dict(dudt_nodes or {})


class Conditions:  # conditions.jl:310-324, 486-552
    def __init__(self, mesh, bc: BoundaryConditions, ic: Optional[InternalConditions] = None):
        ic = ic or InternalConditions()
        self.neumann_edges: Dict[Tuple[int, int], int] = {}
        self.constrained_edges: Dict[Tuple[int, int], int] = {}
        self.dirichlet_nodes: Dict[int, int] = dict(ic.dirichlet_nodes)  # :491
        self.dudt_nodes: Dict[int, int] = dict(ic.dudt_nodes)  # :492
        self.functions = tuple(ic.functions) + tuple(bc.functions)  # :501
        nif = len(ic.functions)
        tri = mesh.triangulation
        # merge_conditions!, conditions.jl:506-544
        for s, nodes in enumerate(tri.boundary_sections):
            ctype = bc.condition_types[s]
            fidx = s + nif  # :518 internal functions first
            for e in range(len(nodes) - 1):
                u, v = int(nodes[e]), int(nodes[e + 1])
                if ctype == Neumann:
                    self.neumann_edges[(u, v)] = fidx
                elif ctype == Constrained:
                    self.constrained_edges[(u, v)] = fidx
                elif ctype == Dirichlet:
                    self.dirichlet_nodes[u] = fidx
                    self.dirichlet_nodes[v] = fidx
                else:
                    self.dudt_nodes[u] = fidx
                    self.dudt_nodes[v] = fidx

    # predicates, conditions.jl:342-484
    def is_dudt_node(self, i):
        return i in self.dudt_nodes

    def is_dirichlet_node(self, i):
        return i in self.dirichlet_nodes

    def is_neumann_edge(self, i, j):
        return (i, j) in self.neumann_edges

    def has_condition(self, i):
        return i in self.dudt_nodes or i in self.dirichlet_nodes

    def eval_condition_fnc(self, fidx, x, y, t, u):
        return self.functions[fidx](x, y, t, u)


# --------------------------------------------------------------------------
# Problems  (/root/reference/src/problem.jl)
# --------------------------------------------------------------------------
def construct_flux_function(q, D, Dp):  # problem.jl:425-440
    if q is None:
        def flux(x, y, t, alpha, beta, gamma, p):
            u = alpha * x + beta * y + gamma
            Dval = D(x, y, t, u, Dp)
            return (-Dval * alpha, -Dval * beta)
        return flux
    return q


def _zero_source(x, y, t, u, p):
    """(x, y, t, u, p) -> zero(eltype(u)); systems pass u as a tuple."""
    return 0.0 * (u[0] if isinstance(u, tuple) else u)


class FVMProblem:  # problem.jl:96-163
    def __init__(self, mesh, boundary_conditions, internal_conditions=None, *,
                 diffusion_function=None, diffusion_parameters=None,
                 source_function=None, source_parameters=None,
                 flux_function=None, flux_parameters=None,
                 initial_condition, initial_time=0.0, final_time):
        ic = np.asarray(initial_condition, dtype=np.float64)
        assert len(ic) == mesh.triangulation.num_points, \
            "The initial condition must have the same number of elements as the number of nodes in the mesh"
        self.mesh = mesh
        if isinstance(boundary_conditions, Conditions):
            self.conditions = boundary_conditions
        else:
            self.conditions = Conditions(mesh, boundary_conditions, internal_conditions)
        self.flux_function = construct_flux_function(flux_function, diffusion_function, diffusion_parameters)
        self.flux_parameters = flux_parameters
        self.source_function = source_function or _zero_source  # problem.jl:123
        self.source_parameters = source_parameters
        self.initial_condition = ic
        self.initial_time = initial_time
        self.final_time = final_time
        self.neqs = 0

    def eval_flux_function(self, x, y, t, a, b, g):  # problem.jl:113-116
        return self.flux_function(x, y, t, a, b, g, self.flux_parameters)

    def eval_source_fnc(self, x, y, t, u):  # problem.jl:10-13
        return self.source_function(x, y, t, u, self.source_parameters)


class FVMSystem:  # problem.jl:233-279, 359-411
    def __init__(self, *probs: FVMProblem):
        assert len(probs) > 0, "There must be at least one problem."
        self.problems = probs
        self.mesh = probs[0].mesh
        assert all(p.mesh is self.mesh for p in probs), "All problems must have the same mesh."
        self.initial_time = probs[0].initial_time
        self.final_time = probs[0].final_time
        self.neqs = len(probs)
        self.initial_condition = np.stack([p.initial_condition for p in probs], axis=1)  # (N, neq)
        self.conditions = tuple(p.conditions for p in probs)  # SimpleConditions per variable
        # every variable keeps its own function tuple; map_fidx (problem.jl:322) is
        # therefore the identity on the per-variable tuple here.

    def eval_flux_function(self, x, y, t, a, b, g):  # problem.jl:354-357, utils.jl:61-70
        return tuple(p.flux_function(x, y, t, a, b, g, p.flux_parameters) for p in self.problems)

    def eval_source_fnc(self, var, x, y, t, u):  # problem.jl:342-345
        p = self.problems[var]
        return p.source_function(x, y, t, u, p.source_parameters)


class SteadyFVMProblem:  # problem.jl:174-180
    def __init__(self, prob):
        self.problem = prob
        self.neqs = prob.neqs


# --------------------------------------------------------------------------
# fvm_eqs!  (/root/reference/src/equations/*.jl), serial order
# --------------------------------------------------------------------------
def get_shape_function_coefficients(mesh, t, T, u, neqs=0):  # shape_functions.jl:2-19
    i, j, k = T
    s = mesh.s[t]
    if neqs == 0:
        a = s[0] * u[i] + s[1] * u[j] + s[2] * u[k]
        b = s[3] * u[i] + s[4] * u[j] + s[5] * u[k]
        g = s[6] * u[i] + s[7] * u[j] + s[8] * u[k]
        return a, b, g
    a = tuple(s[0] * u[i, l] + s[1] * u[j, l] + s[2] * u[k, l] for l in range(neqs))
    b = tuple(s[3] * u[i, l] + s[4] * u[j, l] + s[5] * u[k, l] for l in range(neqs))
    g = tuple(s[6] * u[i, l] + s[7] * u[j, l] + s[8] * u[k, l] for l in range(neqs))
    return a, b, g


def get_flux(prob, t_idx, a, b, g, t, e):  # individual_flux_contributions.jl:27-50
    x, y, nx, ny, l = prob.mesh.get_cv_components(t_idx, e)
    if prob.neqs == 0:
        qx, qy = prob.eval_flux_function(x, y, t, a, b, g)
        return (qx * nx + qy * ny) * l
    q = prob.eval_flux_function(x, y, t, a, b, g)
    return tuple((q[v][0] * nx + q[v][1] * ny) * l for v in range(prob.neqs))


def fvm_eqs_single_triangle(du, u, prob, t, t_idx):  # triangle_contributions.jl:10-35
    T = tuple(int(v) for v in prob.mesh.triangulation.triangles[t_idx])
    i, j, k = T
    a, b, g = get_shape_function_coefficients(prob.mesh, t_idx, T, u, prob.neqs)
    q1 = get_flux(prob, t_idx, a, b, g, t, 0)
    q2 = get_flux(prob, t_idx, a, b, g, t, 1)
    q3 = get_flux(prob, t_idx, a, b, g, t, 2)
    if prob.neqs == 0:
        du[i] = du[i] + q3 - q1
        du[j] = du[j] + q1 - q2
        du[k] = du[k] + q2 - q3
    else:
        for v in range(prob.neqs):
            du[i, v] = du[i, v] + q3[v] - q1[v]
            du[j, v] = du[j, v] + q1[v] - q2[v]
            du[k, v] = du[k, v] + q2[v] - q3[v]


def get_boundary_fluxes(prob, a, b, g, i, j, t):  # boundary_edge_contributions.jl:2-58
    nx, ny, mix, miy, mjx, mjy, l, _, _ = prob.mesh.get_boundary_cv_components(i, j)
    if prob.neqs == 0:
        conds = prob.conditions

        def bflux(x, y):
            ushape = a * x + b * y + g
            if not conds.is_neumann_edge(i, j):
                qx, qy = prob.eval_flux_function(x, y, t, a, b, g)
                return qx * nx + qy * ny
            return conds.eval_condition_fnc(conds.neumann_edges[(i, j)], x, y, t, ushape)

        return bflux(mix, miy) * l, bflux(mjx, mjy) * l
    neqs = prob.neqs

    def bfluxes(x, y):
        ushape = tuple(a[v] * x + b[v] * y + g[v] for v in range(neqs))
        allflux = prob.eval_flux_function(x, y, t, a, b, g)
        out = []
        for v in range(neqs):
            conds = prob.conditions[v]
            if not conds.is_neumann_edge(i, j):
                out.append((allflux[v][0] * nx + allflux[v][1] * ny) * l)
            else:
                out.append(conds.eval_condition_fnc(conds.neumann_edges[(i, j)], x, y, t, ushape) * l)
        return tuple(out)

    return bfluxes(mix, miy), bfluxes(mjx, mjy)


def fvm_eqs_single_boundary_edge(du, u, prob, t, e):  # boundary_edge_contributions.jl:77-86
    i, j = e
    k = prob.mesh.triangulation.get_adjacent(i, j)
    T, t_idx = prob.mesh.safe_get_triangle_props((i, j, k))
    a, b, g = get_shape_function_coefficients(prob.mesh, t_idx, T, u, prob.neqs)
    s1, s2 = get_boundary_fluxes(prob, a, b, g, i, j, t)
    if prob.neqs == 0:
        du[i] = du[i] - s1
        du[j] = du[j] - s2
    else:
        for v in range(prob.neqs):
            du[i, v] = du[i, v] - s1[v]
            du[j, v] = du[j, v] - s2[v]


def fvm_eqs_single_source_contribution(du, u, prob, t, i):  # source_contributions.jl:2-68
    tri = prob.mesh.triangulation
    if prob.neqs == 0:
        if not tri.has_vertex(i):
            du[i] = 0.0
            return
        x, y = tri.points[i]
        conds = prob.conditions
        if not conds.has_condition(i):
            S = prob.eval_source_fnc(x, y, t, u[i])
            du[i] = du[i] / prob.mesh.cv_volumes[i] + S
        elif conds.is_dirichlet_node(i):
            du[i] = 0.0
        else:
            du[i] = conds.eval_condition_fnc(conds.dudt_nodes[i], x, y, t, u[i])
        return
    if not tri.has_vertex(i):
        du[i, :] = 0.0
        return
    x, y = tri.points[i]
    for v in range(prob.neqs):
        conds = prob.conditions[v]
        if not conds.has_condition(i):
            S = prob.eval_source_fnc(v, x, y, t, tuple(u[i]))
            du[i, v] = du[i, v] / prob.mesh.cv_volumes[i] + S
        elif conds.is_dirichlet_node(i):
            du[i, v] = 0.0
        else:
            du[i, v] = conds.eval_condition_fnc(conds.dudt_nodes[i], x, y, t, tuple(u[i]))


def fvm_eqs(du, u, prob, t):
    """serial_fvm_eqs!, main_equations.jl:38-44: zero, triangles, boundary edges, nodes."""
    du[...] = 0.0
    tri = prob.mesh.triangulation
    for t_idx in range(len(tri.triangles)):
        fvm_eqs_single_triangle(du, u, prob, t, t_idx)
    for e in tri.boundary_edge_map.keys():
        fvm_eqs_single_boundary_edge(du, u, prob, t, e)
    for i in range(tri.num_points):
        fvm_eqs_single_source_contribution(du, u, prob, t, i)
    return du


def fvm_eqs_vec(du, u, prob, t):
    """Same arithmetic as :func:`fvm_eqs`, elementwise-vectorised over triangles
    and nodes; needs flux/source/condition callables that broadcast over NumPy
    arrays.  ``np.add.at`` accumulates sequentially in triangle order, so the
    result is bit-identical to the loop version (checked in tests)."""
    mesh = prob.mesh
    tri = mesh.triangulation
    Tr = tri.triangles
    i, j, k = Tr[:, 0], Tr[:, 1], Tr[:, 2]
    s = mesh.s
    neqs = prob.neqs
    du[...] = 0.0
    if neqs == 0:
        a = s[:, 0] * u[i] + s[:, 1] * u[j] + s[:, 2] * u[k]
        b = s[:, 3] * u[i] + s[:, 4] * u[j] + s[:, 5] * u[k]
        g = s[:, 6] * u[i] + s[:, 7] * u[j] + s[:, 8] * u[k]
        Q = []
        for e in range(3):
            qx, qy = prob.eval_flux_function(mesh.mid[:, e, 0], mesh.mid[:, e, 1], t, a, b, g)
            Q.append((qx * mesh.nrm[:, e, 0] + qy * mesh.nrm[:, e, 1]) * mesh.len[:, e])
        q1, q2, q3 = Q
        # per triangle: du[i] = du[i] + q3 - q1 (left to right), in triangle order
        _seq_update(du, Tr, (q3, q1, q2), (q1, q2, q3))
    else:
        a = tuple(s[:, 0] * u[i, l] + s[:, 1] * u[j, l] + s[:, 2] * u[k, l] for l in range(neqs))
        b = tuple(s[:, 3] * u[i, l] + s[:, 4] * u[j, l] + s[:, 5] * u[k, l] for l in range(neqs))
        g = tuple(s[:, 6] * u[i, l] + s[:, 7] * u[j, l] + s[:, 8] * u[k, l] for l in range(neqs))
        Q = []
        for e in range(3):
            q = prob.eval_flux_function(mesh.mid[:, e, 0], mesh.mid[:, e, 1], t, a, b, g)
            Q.append(tuple((q[v][0] * mesh.nrm[:, e, 0] + q[v][1] * mesh.nrm[:, e, 1]) * mesh.len[:, e]
                           for v in range(neqs)))
        for v in range(neqs):
            q1, q2, q3 = Q[0][v], Q[1][v], Q[2][v]
            _seq_update(du[:, v], Tr, (q3, q1, q2), (q1, q2, q3))
    for e in tri.boundary_edge_map.keys():
        fvm_eqs_single_boundary_edge(du, u, prob, t, e)
    _node_pass_vec(du, u, prob, t)
    return du


def _seq_update(du, Tr, plus, minus):
    """du[v] = (du[v] + plus) - minus per triangle vertex, sequential in triangle
    order.  (a + p) - m is not (a + (p - m)); emulate the two roundings with a
    tight loop over the (few) triangles sharing a node: process triangles in
    'rounds' where no node repeats inside a round."""
    T = len(Tr)
    flat_nodes = Tr.ravel()
    P = np.stack(plus, axis=1).ravel()
    M = np.stack(minus, axis=1).ravel()
    # occurrence rank of each (triangle,vertex) entry among entries of the same node
    order = np.argsort(flat_nodes, kind="stable")
    sorted_nodes = flat_nodes[order]
    first = np.r_[True, sorted_nodes[1:] != sorted_nodes[:-1]]
    grp_start = np.maximum.accumulate(np.where(first, np.arange(3 * T), 0))
    rank_sorted = np.arange(3 * T) - grp_start
    rank = np.empty(3 * T, dtype=np.int64)
    rank[order] = rank_sorted
    for r in range(int(rank.max()) + 1 if T else 0):
        sel = np.nonzero(rank == r)[0]
        n = flat_nodes[sel]
        du[n] = du[n] + P[sel] - M[sel]


def _node_pass_vec(du, u, prob, t):
    tri = prob.mesh.triangulation
    N = tri.num_points
    V = prob.mesh.cv_volumes
    x, y = tri.points[:, 0], tri.points[:, 1]
    if prob.neqs == 0:
        special = set(prob.conditions.dirichlet_nodes) | set(prob.conditions.dudt_nodes)
        special |= set(np.nonzero(~tri._vertices)[0].tolist())
        free = np.ones(N, dtype=bool)
        if special:
            free[list(special)] = False
        S = prob.eval_source_fnc(x[free], y[free], t, u[free])
        du[free] = du[free] / V[free] + S
        for i in special:
            fvm_eqs_single_source_contribution(du, u, prob, t, i)
        return
    for v in range(prob.neqs):
        conds = prob.conditions[v]
        special = set(conds.dirichlet_nodes) | set(conds.dudt_nodes)
        special |= set(np.nonzero(~tri._vertices)[0].tolist())
        free = np.ones(N, dtype=bool)
        if special:
            free[list(special)] = False
        uf = tuple(u[free, l] for l in range(prob.neqs))
        S = prob.eval_source_fnc(v, x[free], y[free], t, uf)
        du[free, v] = du[free, v] / V[free] + S
        for i in special:
            if not tri.has_vertex(i):
                du[i, v] = 0.0
                continue
            xi, yi = tri.points[i]
            if conds.is_dirichlet_node(i):
                du[i, v] = 0.0
            else:
                du[i, v] = conds.eval_condition_fnc(conds.dudt_nodes[i], xi, yi, t, tuple(u[i]))


def update_dirichlet_nodes(u, t, prob):  # dirichlet.jl:2-53 (serial)
    pts = prob.mesh.triangulation.points
    if prob.neqs == 0:
        for i, fidx in prob.conditions.dirichlet_nodes.items():
            x, y = pts[i]
            u[i] = prob.conditions.eval_condition_fnc(fidx, x, y, t, u[i])
        return u
    for v in range(prob.neqs):
        conds = prob.conditions[v]
        for i, fidx in conds.dirichlet_nodes.items():
            x, y = pts[i]
            u[i, v] = conds.eval_condition_fnc(fidx, x, y, t, tuple(u[i]))
    return u


def jacobian_sparsity(tri: Triangulation, neqs=0):
    """solve.jl:56-77 (scalar) / :96-131 (system, node-major interleaving).
    Returns sorted (rows, cols) of the structural pattern."""
    nb = tri.get_neighbours()
    rows, cols = [], []
    for i in range(tri.num_points):
        rows.append(i)
        cols.append(i)
        if not tri.has_vertex(i):
            continue
        for j in sorted(nb[i]):
            rows.append(i)
            cols.append(j)
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    if neqs <= 1:
        return rows, cols
    R, C = [], []
    for l in range(neqs):
        for m in range(neqs):
            R.append(rows * neqs + l)
            C.append(cols * neqs + m)
    return np.concatenate(R), np.concatenate(C)


# --------------------------------------------------------------------------
# Templates  (/root/reference/src/specific_problems/*.jl)
# --------------------------------------------------------------------------
class _Acc:
    """Sparse stand-in for the reference's dense ``zeros(n,n)``:
    a dict accumulator with the same ``+=`` order as the reference loops."""

    def __init__(self, n):
        self.n = n
        self.d: Dict[Tuple[int, int], float] = {}

    def add(self, r, c, v):
        self.d[(r, c)] = self.d.get((r, c), 0.0) + v

    def set(self, r, c, v):
        self.d[(r, c)] = v

    def tocsr(self, drop_zeros=True):
        import scipy.sparse as sp
        if not self.d:
            return sp.csr_matrix((self.n, self.n))
        keys = np.array(list(self.d.keys()), dtype=np.int64)
        vals = np.array(list(self.d.values()), dtype=np.float64)
        if drop_zeros:  # sparse(A) drops exact zeros only (Appendix D-1)
            keep = vals != 0.0
            keys, vals = keys[keep], vals[keep]
        A = sp.csr_matrix((vals, (keys[:, 0], keys[:, 1])), shape=(self.n, self.n))
        A.sort_indices()
        return A


def triangle_contributions(A: _Acc, mesh, conditions, diffusion_function, diffusion_parameters):
    """abstract_templates.jl:73-99"""
    Tr = mesh.triangulation.triangles
    V = mesh.cv_volumes
    for t, (i, j, k) in enumerate(Tr.tolist()):
        ijk = (i, j, k)
        s = mesh.s[t]
        for e, (e1, e2) in enumerate(((i, j), (j, k), (k, i))):
            x, y, nx, ny, l = mesh.get_cv_components(t, e)
            D = diffusion_function(x, y, diffusion_parameters)
            Dl = D * l
            a123 = (Dl * (s[0] * nx + s[3] * ny), Dl * (s[1] * nx + s[4] * ny), Dl * (s[2] * nx + s[5] * ny))
            e1c = conditions.has_condition(e1)
            e2c = conditions.has_condition(e2)
            for v in range(3):
                if not e1c:
                    A.add(e1, ijk[v], a123[v] / V[e1])
                if not e2c:
                    A.add(e2, ijk[v], -(a123[v] / V[e2]))


def non_neumann_boundary_edge_contributions(A: _Acc, mesh, conditions, diffusion_function, diffusion_parameters):
    """abstract_templates.jl:237-267, including the ``/V_i`` on the j row (:262, Appendix D-3)."""
    V = mesh.cv_volumes
    for (i, j) in mesh.triangulation.boundary_edge_map.keys():
        if conditions.is_neumann_edge(i, j):
            continue
        nx, ny, mix, miy, mjx, mjy, l, T, t = mesh.get_boundary_cv_components(i, j)
        s = mesh.s[t]
        Di = diffusion_function(mix, miy, diffusion_parameters)
        Dj = diffusion_function(mjx, mjy, diffusion_parameters)
        ic = conditions.has_condition(i)
        jc = conditions.has_condition(j)
        ai = (Di * l * (s[0] * nx + s[3] * ny), Di * l * (s[1] * nx + s[4] * ny), Di * l * (s[2] * nx + s[5] * ny))
        aj = (Dj * l * (s[0] * nx + s[3] * ny), Dj * l * (s[1] * nx + s[4] * ny), Dj * l * (s[2] * nx + s[5] * ny))
        for v in range(3):
            if not ic:
                A.add(i, T[v], ai[v] / V[i])
            if not jc:
                A.add(j, T[v], aj[v] / V[i])  # sic: V[i]


def neumann_boundary_edge_contributions(b, mesh, conditions, diffusion_function, diffusion_parameters):
    """abstract_templates.jl:175-191"""
    V = mesh.cv_volumes
    for (i, j), fidx in conditions.neumann_edges.items():
        _, _, mix, miy, mjx, mjy, l, _, _ = mesh.get_boundary_cv_components(i, j)
        Di = diffusion_function(mix, miy, diffusion_parameters)
        Dj = diffusion_function(mjx, mjy, diffusion_parameters)
        ai = conditions.eval_condition_fnc(fidx, mix, miy, None, None)
        aj = conditions.eval_condition_fnc(fidx, mjx, mjy, None, None)
        if not conditions.has_condition(i):
            b[i] += Di * ai * l / V[i]
        if not conditions.has_condition(j):
            b[j] += Dj * aj * l / V[j]


def boundary_edge_contributions(A, b, mesh, conditions, D, Dp):  # :149-160
    non_neumann_boundary_edge_contributions(A, mesh, conditions, D, Dp)
    neumann_boundary_edge_contributions(b, mesh, conditions, D, Dp)


def apply_dirichlet_conditions(ic, mesh, conditions):  # :109-117
    P = mesh.triangulation.points
    for i, fidx in conditions.dirichlet_nodes.items():
        ic[i] = conditions.eval_condition_fnc(fidx, P[i, 0], P[i, 1], None, None)


def apply_dudt_conditions(b, mesh, conditions):  # :127-135
    P = mesh.triangulation.points
    for i, fidx in conditions.dudt_nodes.items():
        if not conditions.is_dirichlet_node(i):
            b[i] = conditions.eval_condition_fnc(fidx, P[i, 0], P[i, 1], None, None)


def apply_steady_dirichlet_conditions(A, b, mesh, conditions):  # :302-309
    P = mesh.triangulation.points
    for i, fidx in conditions.dirichlet_nodes.items():
        b[i] = conditions.eval_condition_fnc(fidx, P[i, 0], P[i, 1], None, None)
        A.set(i, i, 1.0)


def create_rhs_b(mesh, conditions, source_function, source_parameters):  # :278-288
    tri = mesh.triangulation
    b = np.zeros(tri.num_points)
    for i in tri.each_solid_vertex():
        if not conditions.is_dirichlet_node(int(i)):
            b[i] = source_function(tri.points[i, 0], tri.points[i, 1], source_parameters)
    return b


def fix_missing_vertices(A, b, mesh):  # :317-325
    tri = mesh.triangulation
    for i in range(tri.num_points):
        if not tri.has_vertex(i):
            A.set(i, i, 1.0)
            b[i] = 0.0


@dataclass
class TemplateResult:
    A: object  # scipy csr
    b: np.ndarray
    u0: Optional[np.ndarray] = None
    conditions: object = None
    kind: str = ""


def DiffusionEquation(mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None,
                      initial_condition, initial_time=0.0, final_time):
    """diffusion_equation.jl:69-101.  Returns A (n x n), b and the Dirichlet-fixed
    initial condition (without the trailing 1 of the augmented system, Appendix D-2)."""
    conditions = Conditions(mesh, BCs, ICs)
    n = mesh.triangulation.num_points
    A = _Acc(n)
    b = np.zeros(n)
    ic = np.array(initial_condition, dtype=np.float64)
    triangle_contributions(A, mesh, conditions, diffusion_function, diffusion_parameters)
    boundary_edge_contributions(A, b, mesh, conditions, diffusion_function, diffusion_parameters)
    apply_dudt_conditions(b, mesh, conditions)
    apply_dirichlet_conditions(ic, mesh, conditions)
    fix_missing_vertices(A, b, mesh)
    return TemplateResult(A.tocsr(), b, ic, conditions, "diffusion")


def LinearReactionDiffusionEquation(mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None,
                                    source_function, source_parameters=None,
                                    initial_condition, initial_time=0.0, final_time):
    """linear_reaction_diffusion_equations.jl:76-125"""
    conditions = Conditions(mesh, BCs, ICs)
    tri = mesh.triangulation
    n = tri.num_points
    A = _Acc(n)
    b = np.zeros(n)
    ic = np.array(initial_condition, dtype=np.float64)
    triangle_contributions(A, mesh, conditions, diffusion_function, diffusion_parameters)
    boundary_edge_contributions(A, b, mesh, conditions, diffusion_function, diffusion_parameters)
    for i in tri.each_solid_vertex():  # linear_source_contributions!, :115-125
        i = int(i)
        if not conditions.has_condition(i):
            A.add(i, i, source_function(tri.points[i, 0], tri.points[i, 1], source_parameters))
    apply_dudt_conditions(b, mesh, conditions)
    apply_dirichlet_conditions(ic, mesh, conditions)
    fix_missing_vertices(A, b, mesh)
    return TemplateResult(A.tocsr(), b, ic, conditions, "linear_reaction_diffusion")


def PoissonsEquation(mesh, BCs, ICs=None, *, diffusion_function=lambda x, y, p: 1.0, diffusion_parameters=None,
                     source_function, source_parameters=None):
    """poissons_equation.jl:58-88"""
    conditions = Conditions(mesh, BCs, ICs)
    if conditions.dudt_nodes:
        raise ValueError("PoissonsEquation does not support Dudt nodes.")
    n = mesh.triangulation.num_points
    A = _Acc(n)
    b = create_rhs_b(mesh, conditions, source_function, source_parameters)
    triangle_contributions(A, mesh, conditions, diffusion_function, diffusion_parameters)
    boundary_edge_contributions(A, b, mesh, conditions, diffusion_function, diffusion_parameters)
    apply_steady_dirichlet_conditions(A, b, mesh, conditions)
    fix_missing_vertices(A, b, mesh)
    return TemplateResult(A.tocsr(), b, None, conditions, "poisson")


def LaplacesEquation(mesh, BCs, ICs=None, *, diffusion_function=lambda x, y, p: 1.0, diffusion_parameters=None):
    """laplaces_equation.jl:51-78"""
    conditions = Conditions(mesh, BCs, ICs)
    if conditions.dudt_nodes:
        raise ValueError("PoissonsEquation does not support Dudt nodes.")  # sic, reference message
    n = mesh.triangulation.num_points
    A = _Acc(n)
    b = np.zeros(n)
    triangle_contributions(A, mesh, conditions, diffusion_function, diffusion_parameters)
    boundary_edge_contributions(A, b, mesh, conditions, diffusion_function, diffusion_parameters)
    apply_steady_dirichlet_conditions(A, b, mesh, conditions)
    fix_missing_vertices(A, b, mesh)
    return TemplateResult(A.tocsr(), b, None, conditions, "laplace")


def triangle_contributions_vec(mesh, conditions, diffusion_function, diffusion_parameters):
    """abstract_templates.jl:73-99 vectorised over triangles for 10^6-node parity cases: the same 18 products and
    quotients per triangle, returned as COO triplets; duplicates are summed by SciPy instead of by the reference's
    sequential ``+=`` (summation order only).  ``diffusion_function`` must broadcast over arrays."""
    Tr = mesh.triangulation.triangles
    V = mesh.cv_volumes
    n = mesh.triangulation.num_points
    has_cond = np.zeros(n, dtype=bool)
    for d in (conditions.dirichlet_nodes, conditions.dudt_nodes):
        if d:
            has_cond[np.fromiter(d.keys(), dtype=np.int64, count=len(d))] = True
    s = mesh.s
    rows, cols, vals = [], [], []
    for e, (c1, c2) in enumerate(((0, 1), (1, 2), (2, 0))):
        e1, e2 = Tr[:, c1], Tr[:, c2]
        x, y = mesh.mid[:, e, 0], mesh.mid[:, e, 1]
        nx, ny, l = mesh.nrm[:, e, 0], mesh.nrm[:, e, 1], mesh.len[:, e]
        D = diffusion_function(x, y, diffusion_parameters) + 0.0 * x
        Dl = D * l
        k1, k2 = ~has_cond[e1], ~has_cond[e2]
        for v in range(3):
            a = Dl * (s[:, v] * nx + s[:, 3 + v] * ny)
            rows += [e1[k1], e2[k2]]
            cols += [Tr[k1, v], Tr[k2, v]]
            vals += [a[k1] / V[e1[k1]], -(a[k2] / V[e2[k2]])]
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def MeanExitTimeProblem(mesh, BCs, ICs=None, *, diffusion_function, diffusion_parameters=None, vectorised=False):
    """mean_exit_time.jl:57-94: no boundary-edge pass, BC functions never evaluated."""
    conditions = Conditions(mesh, BCs, ICs)
    if conditions.dudt_nodes:
        raise ValueError("MeanExitTimeProblem does not support Dudt nodes.")
    if conditions.constrained_edges:
        raise ValueError("MeanExitTimeProblem does not support Constrained edges.")
    tri = mesh.triangulation
    n = tri.num_points
    if vectorised:
        import scipy.sparse as sp
        r, c, v = triangle_contributions_vec(mesh, conditions, diffusion_function, diffusion_parameters)
        is_vertex = np.asarray(tri._vertices, dtype=bool)
        dirn = np.zeros(n, dtype=bool)
        if conditions.dirichlet_nodes:
            dirn[np.fromiter(conditions.dirichlet_nodes.keys(), dtype=np.int64, count=len(conditions.dirichlet_nodes))] = True
        ident = np.nonzero((dirn & is_vertex) | ~is_vertex)[0]  # create_met_b! :84-94 and fix_missing_vertices :317-325
        A = sp.csr_matrix((np.r_[v, np.ones(len(ident))], (np.r_[r, ident], np.r_[c, ident])), shape=(n, n))
        A.sum_duplicates()
        A.eliminate_zeros()  # sparse(A) drops exact zeros (Appendix D-1)
        A.sort_indices()
        b = np.where(is_vertex & ~dirn, -1.0, 0.0)
        return TemplateResult(A, b, None, conditions, "mean_exit_time")
    A = _Acc(n)
    triangle_contributions(A, mesh, conditions, diffusion_function, diffusion_parameters)
    b = np.zeros(n)
    for i in tri.each_solid_vertex():  # create_met_b!, :84-94
        i = int(i)
        if not conditions.is_dirichlet_node(i):
            b[i] = -1
        else:
            A.set(i, i, 1.0)
    fix_missing_vertices(A, b, mesh)
    return TemplateResult(A.tocsr(), b, None, conditions, "mean_exit_time")


def solve_steady(tpl: TemplateResult):
    """LinearProblem(A,b) solved with a sparse direct method (SciPy SuperLU standing
    in for the reference's KLUFactorization, docs/src/literate_wyos/poissons_equation.jl:101)."""
    import scipy.sparse.linalg as spla
    return spla.spsolve(tpl.A.tocsc(), tpl.b)


# --------------------------------------------------------------------------
# Fixed-step Tsit5 (OrdinaryDiffEq tableau; parity unpinned)
# --------------------------------------------------------------------------
TSIT5_C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
TSIT5_A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)


def tsit5_fixed(f, u0, t0, t1, dt, callback=None, fsal=True, saveat=None):
    """``solve(prob, Tsit5(); adaptive=false, dt)``: k1..k6 stages, u+ = u + dt*sum(a7j kj),
    k7 = f(u+) reused as next k1 (FSAL) unless the callback modified u.  The callback
    fires after every accepted step and not at t0 (Appendix C).
    ``f(du,u,t)`` in place.  Returns u(t1) (and the saved states if saveat)."""
    u = np.array(u0, dtype=np.float64)
    nsteps = int(round((t1 - t0) / dt))
    assert abs(t0 + nsteps * dt - t1) <= 1e-12 * max(1.0, abs(t1)), "dt must divide the time span"
    k = [np.zeros_like(u) for _ in range(7)]
    tmp = np.zeros_like(u)
    have_k1 = False
    saves = []
    t = t0
    for n in range(nsteps):
        t = t0 + n * dt
        if not have_k1:
            f(k[0], u, t)
        for s in range(1, 6):
            tmp[...] = u
            acc = np.zeros_like(u)
            for jj in range(s):
                acc += TSIT5_A[s][jj] * k[jj]
            tmp += dt * acc
            f(k[s], tmp, t + TSIT5_C[s] * dt)
        acc = np.zeros_like(u)
        for jj in range(6):
            acc += TSIT5_A[6][jj] * k[jj]
        u = u + dt * acc
        tn = t0 + (n + 1) * dt
        modified = False
        if callback is not None:
            modified = bool(callback(u, tn))
        if fsal and not modified:
            f(k[6], u, tn)
            k[0], k[6] = k[6], k[0]
            have_k1 = True
        else:
            have_k1 = False
        if saveat is not None and any(abs(tn - ts) < 0.5 * dt for ts in saveat):
            saves.append(u.copy())
    if saveat is not None:
        return u, saves
    return u


TSIT5_BTILDE = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                0.5823571654525552, -0.45808210592918697, 0.015151515151515152)


def tsit5_adaptive(f, u0, t0, t1, abstol=1e-6, reltol=1e-3, dt0=None, callback=None, saveat=None):
    """Adaptive Tsit5 with the embedded error estimate, OrdinaryDiffEq's scaled RMS norm and PI
    controller (beta1 = 7/50, beta2 = 2/25, gamma = 0.9, qmin = 0.2, qmax = 10, qsteady in [1, 1.2]),
    Hairer's initial step, saveat times as tstops.  The controller lives in OrdinaryDiffEq (absent
    from /root/reference): parity unpinned; this restates the documented algorithm.
    Returns (u(t1), saves, n_accept, n_reject)."""
    u = np.array(u0, dtype=np.float64)
    n = u.size
    k = [np.zeros_like(u) for _ in range(7)]
    t = t0
    saves = []
    saveat = list(saveat or [])
    nxt = 0
    eps_t = 1e-12 * max(1.0, abs(t1))

    def do_save():
        nonlocal nxt
        while nxt < len(saveat) and abs(saveat[nxt] - t) <= 1e-12 * max(1.0, abs(t)):
            saves.append(u.copy())
            nxt += 1

    do_save()
    f(k[0], u, t)
    if dt0 is None or dt0 <= 0:
        sc = abstol + np.abs(u) * reltol
        d0 = math.sqrt(np.sum((u / sc) ** 2) / n)
        d1 = math.sqrt(np.sum((k[0] / sc) ** 2) / n)
        dt = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        dt = min(dt, t1 - t0)
    else:
        dt = dt0
    beta1, beta2, gamma, qmin, qmax = 7 / 50, 2 / 25, 0.9, 0.2, 10.0
    qold = 1e-4
    nacc = nrej = 0
    have_k1 = True
    while t < t1 - eps_t:
        tstop = t1
        if nxt < len(saveat) and saveat[nxt] > t + eps_t:
            tstop = min(tstop, saveat[nxt])
        clipped = False
        h = dt
        if t + h >= tstop - eps_t:
            h = tstop - t
            clipped = True
        if not have_k1:
            f(k[0], u, t)
        have_k1 = True
        for s in range(1, 6):
            acc = TSIT5_A[s][0] * k[0]
            for j in range(1, s):
                acc = acc + TSIT5_A[s][j] * k[j]
            f(k[s], u + h * acc, t + TSIT5_C[s] * h)
        acc = TSIT5_A[6][0] * k[0]
        for j in range(1, 6):
            acc = acc + TSIT5_A[6][j] * k[j]
        unew = u + h * acc
        f(k[6], unew, t + h)
        e = TSIT5_BTILDE[0] * k[0]
        for j in range(1, 7):
            e = e + TSIT5_BTILDE[j] * k[j]
        e = e * h
        sc = abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol
        EEst = math.sqrt(np.sum((e / sc) ** 2) / n)
        if EEst == 0.0:
            q = 1.0 / qmax
        else:
            q = EEst ** beta1 / qold ** beta2
            q = max(1.0 / qmax, min(1.0 / qmin, q / gamma))
        if EEst <= 1.0:
            nacc += 1
            t = tstop if clipped else t + h
            u = unew
            qold = max(EEst, 1e-4)
            dt_new = h / q
            if h <= dt_new <= 1.2 * h:
                dt_new = h
            if (not clipped) or dt_new < dt:
                dt = dt_new
            if callback is not None and callback(u, t):
                have_k1 = False
            else:
                k[0], k[6] = k[6], k[0]
            do_save()
        else:
            nrej += 1
            dt = h / min(1.0 / qmin, q)
    return u, saves, nacc, nrej


def pl_interpolate(prob, T, u, x, y):  # utils.jl:23-27
    Ts, t_idx = prob.mesh.safe_get_triangle_props(tuple(int(v) for v in T))
    a, b, g = get_shape_function_coefficients(prob.mesh, t_idx, Ts, u, prob.neqs)
    if prob.neqs == 0:
        return a * x + b * y + g
    return tuple(a[v] * x + b[v] * y + g[v] for v in range(prob.neqs))


def compute_flux(prob, i, j, u, t):  # problem.jl:458-487
    tri = prob.mesh.triangulation
    px, py = tri.points[i]
    qx, qy = tri.points[j]
    ex, ey = qx - px, qy - py
    ell = math.sqrt(ex * ex + ey * ey)
    nx, ny = ey / ell, -ex / ell
    k = tri.adjacent.get((j, i), -1)
    if k < 0:
        k = tri.get_adjacent(i, j)
    else:
        i, j = j, i
    Ts, t_idx = prob.mesh.safe_get_triangle_props((i, j, k))
    a, b, g = get_shape_function_coefficients(prob.mesh, t_idx, Ts, u, prob.neqs)
    mx, my = (px + qx) / 2, (py + qy) / 2
    qv = prob.eval_flux_function(mx, my, t, a, b, g)
    if prob.neqs == 0:
        return nx * qv[0] + ny * qv[1]
    return tuple(nx * qv[v][0] + ny * qv[v][1] for v in range(prob.neqs))
