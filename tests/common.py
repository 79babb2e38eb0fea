"""Builders that construct the SAME problem twice: for the CPU oracle (closures, like the
reference) and for the GPU host API (registry specs).  Test infrastructure only."""
import math

import numpy as np

import fvm_b200 as G
from oracle import fvm_oracle as O

RTOL_RHS = 1e-12   # BASELINE.json: single RHS evaluation, fp64, only summation order differs
RTOL_TSIT5 = 1e-10  # BASELINE.json: fixed-step solution at final time


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


# ---- registry spec -> oracle closure (NumPy-broadcastable) -------------------------------------
def cond_closure(spec):
    if isinstance(spec, (int, float)):
        spec = G.Const(float(spec))
    if isinstance(spec, G.Const):
        return lambda x, y, t, u, p: spec.c + 0.0 * x
    if isinstance(spec, G.AffineU):
        return lambda x, y, t, u, p: spec.c0 + spec.c1 * (u[p] if isinstance(u, tuple) else u)
    if isinstance(spec, G.ExpSaturation):
        return lambda x, y, t, u, p: spec.c0 * (1.0 - math.exp(-t / spec.tau)) + 0.0 * x
    if isinstance(spec, G.LinearXY):
        return lambda x, y, t, u, p: spec.c0 + spec.cx * x + spec.cy * y
    if isinstance(spec, G.ExpXYT):
        return lambda x, y, t, u, p: spec.c0 * np.exp(spec.cx * x + spec.cy * y + spec.ct * t)
    raise TypeError(spec)


def flux_closure(spec, var=0):
    """returns (flux_function or None, diffusion_function or None)"""
    if isinstance(spec, G.ConstantDiffusion):
        if var is None:
            return None, (lambda x, y, t, u, p: spec.D)
        return (lambda x, y, t, a, b, g, p: (-spec.D * a[var], -spec.D * b[var])), None
    if isinstance(spec, G.TabulatedDiffusion):
        assert var is None
        return None, (lambda x, y, t, u, p: spec.fn(x, y))
    if isinstance(spec, G.PowerDiffusion):
        def Dfun(x, y, t, u, p):
            base = np.abs(u) if spec.use_abs else u
            return spec.D0 * base ** (spec.m - 1)
        if var is None:
            return None, Dfun

        def q(x, y, t, a, b, g, p):
            u = a[var] * x + b[var] * y + g[var]
            D = Dfun(x, y, t, u, p)
            return (-D * a[var], -D * b[var])
        return q, None
    if isinstance(spec, G.AdvectionDiffusionFlux):
        def q(x, y, t, a, b, g, p):
            if var is not None:
                a, b, g = a[var], b[var], g[var]
            u = a * x + b * y + g
            return (spec.nu_x * u - spec.D * a, spec.nu_y * u - spec.D * b)
        return q, None
    if isinstance(spec, G.KellerSegelFlux):
        if var == 0:
            def q(x, y, t, a, b, g, p):  # src/FiniteVolumeMethod.jl:98-104
                u = a[0] * x + b[0] * y + g[0]
                chi = spec.c * u / (1 + u**2)
                return (chi * a[1] - a[0], chi * b[1] - b[0])
            return q, None
        return (lambda x, y, t, a, b, g, p: (-spec.D * a[1], -spec.D * b[1])), None
    raise TypeError(spec)


def source_closure(spec, var=None):
    if spec is None or isinstance(spec, G.ZeroSource):
        return None
    pick = (lambda u: u) if var is None else (lambda u: u[var])
    if isinstance(spec, G.LinearSource):
        return lambda x, y, t, u, p: spec.lam * pick(u) + spec.mu
    if isinstance(spec, G.LogisticSource):
        return lambda x, y, t, u, p: spec.lam * pick(u) * (1 - pick(u))
    if isinstance(spec, G.TabulatedSource):
        return lambda x, y, t, u, p: spec.fn(x, y)
    if isinstance(spec, G.GrayScottSource):
        if var == 0:
            return lambda x, y, t, u, p: spec.b * (1 - u[0]) - u[0] * u[1]**2
        return lambda x, y, t, u, p: -spec.d * u[1] + u[0] * u[1]**2
    if isinstance(spec, G.BrusselatorSource):
        if var == 0:
            return lambda x, y, t, u, p: u[0]**2 * u[1] - 2 * u[0]
        return lambda x, y, t, u, p: -u[0]**2 * u[1] + u[0]
    if isinstance(spec, G.KellerSegelSource):
        if var == 0:
            return lambda x, y, t, u, p: u[0] * (1 - u[0])
        return lambda x, y, t, u, p: u[0] - spec.a * u[1]
    raise TypeError(spec)


def to_oracle_tri(tri):
    return O.Triangulation(tri.points, tri.triangles.astype(np.int64), [np.asarray(s) for s in tri.boundary_sections])


class Pair:
    """One mesh, built for both sides."""

    def __init__(self, gtri):
        self.gtri = gtri
        self.gmesh = G.FVMGeometry(gtri)
        self.otri = to_oracle_tri(gtri)
        self.omesh = O.FVMGeometry(self.otri)

    def problem(self, bc_specs, bc_types, flux, source=None, ic=None, internal=None, var=None, final_time=1.0):
        """scalar problem (var=None) or one member of a system (var = species index)."""
        N = self.gtri.num_points
        if ic is None:
            ic = np.zeros(N)
        if not isinstance(bc_specs, (tuple, list)):
            bc_specs, bc_types = (bc_specs,), (bc_types,)
        gBC = G.BoundaryConditions(self.gmesh, tuple(bc_specs), tuple(bc_types))
        oBC = O.BoundaryConditions(self.omesh, tuple(cond_closure(s) for s in bc_specs), tuple(bc_types),
                                   parameters=(var,) * len(bc_specs))
        gIC = oIC = None
        if internal is not None:
            specs, dnodes, tnodes = internal
            gIC = G.InternalConditions(tuple(specs), dirichlet_nodes=dnodes, dudt_nodes=tnodes)
            oIC = O.InternalConditions(tuple(cond_closure(s) for s in specs), dirichlet_nodes=dnodes, dudt_nodes=tnodes,
                                       parameters=(var,) * len(specs))
        q, D = flux_closure(flux, var)
        S = source_closure(source, var)
        if D is not None:
            gp = G.FVMProblem(self.gmesh, gBC, gIC, diffusion_function=flux, source_function=source,
                              initial_condition=ic, final_time=final_time)
            op = O.FVMProblem(self.omesh, oBC, oIC, diffusion_function=D, source_function=S,
                              initial_condition=ic, final_time=final_time)
        else:
            gp = G.FVMProblem(self.gmesh, gBC, gIC, flux_function=flux, source_function=source,
                              initial_condition=ic, final_time=final_time)
            op = O.FVMProblem(self.omesh, oBC, oIC, flux_function=q, source_function=S,
                              initial_condition=ic, final_time=final_time)
        return gp, op


def delaunay_mesh(n_points, seed, extra_points=0, jitter=None):
    """Unstructured test mesh: Delaunay triangulation of points in the unit square plus a boundary ring;
    optionally trailing points that are not vertices.  jitter=None: uniformly random interior points
    (contains thin triangles); jitter=a: a lattice whose interior nodes are displaced by up to a*h, i.e.
    a well-shaped mesh with unstructured connectivity (node degrees 4..8)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    k = max(4, int(math.sqrt(n_points)))
    ring = np.linspace(0, 1, k, endpoint=False)
    bnd = np.concatenate([np.stack([ring, 0 * ring], 1), np.stack([1 + 0 * ring, ring], 1),
                          np.stack([1 - ring, 1 + 0 * ring], 1), np.stack([0 * ring, 1 - ring], 1)])
    if jitter is None:
        inner = 0.02 + 0.96 * rng.random((n_points, 2))
    else:
        h = 1.0 / k
        g = (np.arange(1, k) * h)
        X, Y = np.meshgrid(g, g)
        inner = np.stack([X.ravel(), Y.ravel()], 1) + jitter * h * (2 * rng.random(((k - 1) ** 2, 2)) - 1)
    pts = np.concatenate([bnd, inner])
    dl = Delaunay(pts)
    tris = dl.simplices.astype(np.int32)
    p, q, r = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    area2 = (q[:, 0] - p[:, 0]) * (r[:, 1] - p[:, 1]) - (q[:, 1] - p[:, 1]) * (r[:, 0] - p[:, 0])
    tris[area2 < 0] = tris[area2 < 0][:, [0, 2, 1]]  # make every triangle ccw
    tris = tris[np.abs(area2) > 1e-12]
    if extra_points:
        pts = np.concatenate([pts, 2.0 + rng.random((extra_points, 2))])
    return G.Triangulation(pts, tris)


def annulus_mesh(nr, nt, r_in=0.2, r_out=1.0):
    """Polar-grid triangulation of an annulus (the reference's tests and the 'diffusion equation on an
    annulus' tutorial use annuli): two boundary loops, outer counter-clockwise and inner clockwise, so the
    interior is on the left of both.  Sections are ordered by their smallest node id: inner first."""
    r = np.linspace(r_in, r_out, nr)
    th = np.linspace(0, 2 * np.pi, nt, endpoint=False)
    pts = np.stack([np.outer(r, np.cos(th)).ravel(), np.outer(r, np.sin(th)).ravel()], 1)  # node = k*nt + m
    tris = []
    for k in range(nr - 1):
        for m in range(nt):
            a, b = k * nt + m, k * nt + (m + 1) % nt
            c, d = (k + 1) * nt + m, (k + 1) * nt + (m + 1) % nt
            tris.append((a, c, b) if False else (a, c, d))
            tris.append((a, d, b))
    tris = np.asarray(tris, dtype=np.int32)
    p, q, s = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    area2 = (q[:, 0] - p[:, 0]) * (s[:, 1] - p[:, 1]) - (q[:, 1] - p[:, 1]) * (s[:, 0] - p[:, 0])
    tris[area2 < 0] = tris[area2 < 0][:, [0, 2, 1]]
    return G.Triangulation(pts, tris)


def disk_mesh(nr, radius=1.0):
    """Disk of the 'reaction-diffusion equation with a time-dependent Dirichlet boundary condition on a disk' tutorial:
    centre node, ring k = 1..nr with 6k nodes, Delaunay-triangulated (the domain is convex); one boundary loop."""
    from scipy.spatial import Delaunay
    pts = [(0.0, 0.0)]
    for k in range(1, nr + 1):
        th = 2 * np.pi * (np.arange(6 * k) + 0.5 * (k % 2)) / (6 * k)
        pts += list(zip(radius * k / nr * np.cos(th), radius * k / nr * np.sin(th)))
    pts = np.asarray(pts)
    tris = Delaunay(pts).simplices.astype(np.int32)
    p, q, r = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    area2 = (q[:, 0] - p[:, 0]) * (r[:, 1] - p[:, 1]) - (q[:, 1] - p[:, 1]) * (r[:, 0] - p[:, 0])
    tris[area2 < 0] = tris[area2 < 0][:, [0, 2, 1]]
    return G.Triangulation(pts, tris[np.abs(area2) > 1e-12])


def wedge_mesh(nr, alpha=np.pi / 4):
    """Circular wedge 0 <= r <= 1, 0 <= theta <= alpha of the 'diffusion equation in a wedge with mixed boundary
    conditions' tutorial: apex node, ring k = 1..nr with k+1 nodes; three boundary sections in the tutorial's order
    (bottom edge, arc, upper edge), consecutive sections sharing their end node."""
    ring0 = [0]
    pts = [(0.0, 0.0)]
    rings = [ring0]
    for k in range(1, nr + 1):
        th = alpha * np.arange(k + 1) / k
        rings.append(list(range(len(pts), len(pts) + k + 1)))
        pts += list(zip(k / nr * np.cos(th), k / nr * np.sin(th)))
    tris = []
    for k in range(nr):
        a, b = rings[k], rings[k + 1]
        for j in range(k):
            tris.append((a[j], b[j], b[j + 1]))
            tris.append((a[j], b[j + 1], a[j + 1]))
        tris.append((a[k], b[k], b[k + 1]))
    sections = [[r[0] for r in rings], rings[nr], [r[-1] for r in rings[::-1]]]
    return G.Triangulation(np.asarray(pts), np.asarray(tris, dtype=np.int32), boundary_sections=sections)
