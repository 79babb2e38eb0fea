"""The problems behind tests/golden/*.npz — shared by the generator (tests/golden/make_golden.py, CPU
oracle only) and by tests/test_golden.py (oracle vs fixture on the CPU, CUDA path vs fixture on the GPU).

Every case fixes the mesh ARRAYS (stored in the fixture, so the unstructured cases do not depend on the
SciPy Delaunay version), the state `u`, the time `t` and the problem; each mirrors a workload of the
reference's tests / tutorials (citations relative to /root/reference).  Test infrastructure only."""
import numpy as np

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import Pair, cond_closure, delaunay_mesh

SEED = 20240517


def _tri_from(fx, default):
    """The mesh of a case: from the fixture when it exists (bit-exact inputs), else freshly built."""
    if fx is None:
        return default()
    ptr = fx["boundary_ptr"]
    secs = [fx["boundary_nodes"][ptr[s]:ptr[s + 1]] for s in range(len(ptr) - 1)]
    return G.Triangulation(fx["points"], fx["triangles"], secs)


def _split(gtri, k):
    loop = gtri.boundary_sections[0]
    n = len(loop) - 1
    cuts = [round(i * n / k) for i in range(k + 1)]
    return G.Triangulation(gtri.points, gtri.triangles, [loop[cuts[i]:cuts[i + 1] + 1] for i in range(k)])


class Case:
    """pair, (gp, op) problems for both sides, state u (N,) or (N,neq), time t, and optional Tsit5 /
    template data."""
    tsit5 = None      # (dt, t1) fixed-step run from prob.initial_condition
    template = None   # name of the linear template built from the same mesh/conditions


def readme_50x50(fx=None):
    """BASELINE configs[0]: README.md:20-45 / docs/src/literate_tutorials/diffusion_equation_on_a_square_plate.jl."""
    c = Case()
    c.pair = Pair(_tri_from(fx, lambda: G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True)))
    ic = np.where(c.pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
    c.gp, c.op = c.pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9), ic=ic, final_time=0.1)
    c.u, c.t = 50 * np.random.default_rng(SEED).random(len(ic)), 0.0
    c.tsit5 = (0.0025, 0.1)
    c.template = "diffusion"
    return c


def convection_robin_24(fx=None):
    """Advection-diffusion with a linear source on a four-section lattice: Neumann with a u-dependent
    (Robin-like) function, Dudt, Dirichlet and homogeneous Neumann sections (the mix of
    test/test_functions.jl:304-374)."""
    c = Case()
    c.pair = Pair(_tri_from(fx, lambda: G.triangulate_rectangle(0, 1, 0, 2, 24, 31, single_boundary=False)))
    P = c.pair.gtri.points
    ic = 0.3 + 0.2 * np.sin(3 * P[:, 0]) * np.cos(2 * P[:, 1])
    specs = (G.AffineU(0.2, -0.3), G.AffineU(0.0, -0.5), G.LinearXY(0.1, 0.2, -0.1), G.Const(0.0))
    types = (G.Neumann, G.Dudt, G.Dirichlet, G.Neumann)
    c.gp, c.op = c.pair.problem(specs, types, G.AdvectionDiffusionFlux(0.05, 0.4, -0.2), source=G.LinearSource(-0.3, 0.1),
                                ic=ic, final_time=0.05)
    c.u, c.t = 0.2 + np.random.default_rng(SEED + 1).random(len(ic)), 0.4
    c.tsit5 = (0.0005, 0.05)
    return c


def keller_segel_16(fx=None):
    """BASELINE configs[3] model (src/FiniteVolumeMethod.jl:92-138, docs/src/tutorials/keller_segel_chemotaxis.md:19-60)
    on a 16x16 lattice: 2-species FVMSystem, all-Neumann, FSAL Tsit5."""
    c = Case()
    c.pair = Pair(_tri_from(fx, lambda: G.triangulate_rectangle(0, 4, 0, 4, 16, 16, single_boundary=True)))
    N = c.pair.gtri.num_points
    rng = np.random.default_rng(SEED + 2)
    U0 = np.ascontiguousarray(np.stack([0.01 * rng.random(N) + 0.5, 0.1 * rng.random(N)], axis=1))
    flux, src = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    g1, o1 = c.pair.problem(G.Const(0.0), G.Neumann, flux, source=src, var=0, ic=U0[:, 0], final_time=0.2)
    g2, o2 = c.pair.problem(G.Const(0.0), G.Neumann, flux, source=src, var=1, ic=U0[:, 1], final_time=0.2)
    c.gp, c.op = G.FVMSystem(g1, g2), O.FVMSystem(o1, o2)
    c.u, c.t = np.ascontiguousarray(np.stack([0.2 + rng.random(N), rng.random(N)], axis=1)), 0.0
    c.tsit5 = (0.004, 0.2)
    return c


def unstructured_porous(fx=None):
    """Unstructured Delaunay mesh with points that are not vertices; Dirichlet / Dudt / Neumann / Constrained
    sections and internal conditions; porous-medium diffusion D0 |u|^(m-1) with a logistic source
    (docs/src/literate_tutorials/porous_medium_equation.jl:50, porous_fisher_equation_and_travelling_waves.jl:60-61)."""
    c = Case()
    c.pair = Pair(_tri_from(fx, lambda: _split(delaunay_mesh(260, 3, extra_points=3, jitter=0.35), 4)))
    N = c.pair.gtri.num_points
    specs = (G.Const(0.25), G.AffineU(0.1, -0.5), G.LinearXY(0.3, 0.2, -0.1), G.Const(0.0))
    types = (G.Dirichlet, G.Dudt, G.Neumann, G.Constrained)
    internal = ((G.Const(0.7), G.AffineU(0.0, 1.0)), {100: 0, 101: 0}, {150: 1, 100: 1})
    c.gp, c.op = c.pair.problem(specs, types, G.PowerDiffusion(0.3, 2.5, use_abs=True), source=G.LogisticSource(1.3),
                                internal=internal)
    c.u, c.t = 0.2 + np.random.default_rng(SEED + 3).random(N), 0.3
    return c


def mean_exit_time_unstructured(fx=None):
    """MeanExitTimeProblem (src/specific_problems/mean_exit_time.jl:57-94) on an unstructured mesh with a
    reflecting (Neumann) and an absorbing (Dirichlet) section — the steady template path of BASELINE configs[2]."""
    c = Case()
    c.pair = Pair(_tri_from(fx, lambda: _split(delaunay_mesh(400, 8, extra_points=2, jitter=0.3), 2)))
    c.gp = c.op = None
    c.u, c.t = np.random.default_rng(SEED + 4).random(c.pair.gtri.num_points), 0.0
    c.template = "met"
    return c


CASES = {f.__name__: f for f in (readme_50x50, convection_robin_24, keller_segel_16, unstructured_porous,
                                 mean_exit_time_unstructured)}


# ---- both sides of a template case ---------------------------------------------------------------
def build_template(c, side, **kw):
    """side 'oracle' -> oracle TemplateResult; side 'gpu' -> fvm_b200 template object (kw: engine options)."""
    const0 = lambda x, y, t, u, p: 0.0 * x
    if c.template == "diffusion":
        ic = np.where(c.pair.gtri.points[:, 1] <= 1.0, 50.0, 0.0)
        if side == "oracle":
            return O.DiffusionEquation(c.pair.omesh, O.BoundaryConditions(c.pair.omesh, const0, O.Dirichlet),
                                       diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.1)
        return G.DiffusionEquation(c.pair.gmesh, G.BoundaryConditions(c.pair.gmesh, G.Const(0.0), G.Dirichlet),
                                   diffusion_function=1 / 9, initial_condition=ic, final_time=0.1, **kw)
    if c.template == "met":
        if side == "oracle":
            f99 = cond_closure(G.Const(99.0))
            bc = O.BoundaryConditions(c.pair.omesh, (lambda x, y, t, u, p: f99(x, y, 0.0, 0.0, p),) * 2, (O.Neumann, O.Dirichlet))
            return O.MeanExitTimeProblem(c.pair.omesh, bc, diffusion_function=lambda x, y, p: 2.5e-3)
        bc = G.BoundaryConditions(c.pair.gmesh, (G.Const(99.0), G.Const(99.0)), (G.Neumann, G.Dirichlet))
        return G.MeanExitTimeProblem(c.pair.gmesh, bc, diffusion_function=2.5e-3, **kw)
    raise KeyError(c.template)


def oracle_outputs(c):
    """Everything a fixture stores besides the inputs, computed by the CPU oracle."""
    out = {}
    if c.op is not None:
        out["du"] = O.fvm_eqs(np.zeros_like(c.u), c.u, c.op, c.t)  # the loop restatement, reference order
        if c.tsit5:
            dt, t1 = c.tsit5
            has_dir = any(len(p.conditions.dirichlet_nodes) > 0 for p in getattr(c.op, "problems", [c.op]))
            cb = (lambda u, t: (O.update_dirichlet_nodes(u, t, c.op), True)[1]) if has_dir else None
            u0 = np.array(c.op.initial_condition, dtype=np.float64)
            res = O.tsit5_fixed(lambda du, u, t: O.fvm_eqs_vec(du, u, c.op, t), u0, 0.0, t1, dt, callback=cb)
            out["tsit5_end"] = res[0] if isinstance(res, tuple) else res
    if c.template:
        ref = build_template(c, "oracle")
        A = ref.A.tocsr()
        A.sort_indices()
        out["A_indptr"], out["A_indices"], out["A_data"], out["b"] = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, ref.b
        out["Au_b"] = ref.A @ c.u + ref.b
        if c.template == "met":
            out["steady"] = O.solve_steady(ref)
        else:
            out["u0"] = ref.u0
            dt, t1 = c.tsit5

            def f(du, u, t):
                du[...] = ref.A @ u + ref.b
            res = O.tsit5_fixed(f, ref.u0, 0.0, t1, dt)
            out["tsit5_operator_end"] = res[0] if isinstance(res, tuple) else res
    return out
