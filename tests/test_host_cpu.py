"""CPU-side checks: the C-ABI library loads and exports every symbol of include/fvmcuda.h, the host
mirror reproduces the reference's constructors/assertions/condition flattening, integer mesh arrays
are bit-exact against the oracle, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import Pair, delaunay_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fvmcuda.h")).read()
    declared = sorted(set(re.findall(r"\b(fvm_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(G.exported_symbols())
    lib = ctypes.CDLL(G.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.fvm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fvm_version()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "finitevolumemethod.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device failure path")
def test_no_cpu_fallback_without_device():
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 4, 4, single_boundary=True))
    gp, _ = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1.0))
    with pytest.raises(G.FVMCudaError) as e:
        G.get_cuda_parameters(gp)
    assert "no CPU fallback" in str(e.value)


def test_lattice_connectivity_bit_exact():
    """Integer mesh/connectivity arrays bit-exact vs the oracle."""
    for single in (True, False):
        g = G.triangulate_rectangle(0.0, 2.0, -1.0, 3.0, 12, 19, single_boundary=single)
        o = O.triangulate_rectangle(0.0, 2.0, -1.0, 3.0, 12, 19, single_boundary=single)
        assert np.array_equal(g.points, o.points)
        assert np.array_equal(g.triangles, o.triangles)
        assert len(g.boundary_sections) == len(o.boundary_sections)
        for a, b in zip(g.boundary_sections, o.boundary_sections):
            assert np.array_equal(a, b)
        uv, _ = g.boundary_edges()
        assert [tuple(e) for e in uv.tolist()] == list(o.boundary_edge_map.keys())
    # test/test_functions.jl:112: 7+(i-1)*12 is a vertical node line for nx = 12 (1-based)
    assert np.all(g.points[6 + 12 * np.arange(19), 0] == g.points[6, 0])
    # test/equations.jl:42: (1,2,201) is a stored triangle for nx = 200
    assert (G.triangulate_rectangle(0, 1, 0, 1, 200, 200).triangles[0] == [0, 1, 200]).all()


def test_chained_boundary_matches_oracle():
    g = delaunay_mesh(300, 2)
    o = O.Triangulation(g.points, g.triangles.astype(np.int64))
    assert len(g.boundary_sections) == len(o.boundary_sections) == 1
    assert np.array_equal(g.boundary_sections[0], o.boundary_sections[0])


def test_conditions_flattening_matches_reference_dicts():
    """merge_conditions! (conditions.jl:506-544): same edge/node maps as the oracle's Dicts."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 9, 7, single_boundary=False))
    specs = (G.Const(1.0), G.Const(2.0), G.Const(3.0), G.Const(4.0))
    types = (G.Neumann, G.Dirichlet, G.Dudt, G.Constrained)
    internal = ((G.Const(5.0),), {20: 0, 21: 0}, {22: 0, 20: 0})
    gp, op = pair.problem(specs, types, G.ConstantDiffusion(1.0), internal=internal)
    gc, oc = gp.conditions, op.conditions
    assert gc.get_dirichlet_nodes() == dict(sorted(oc.dirichlet_nodes.items()))
    assert gc.get_dudt_nodes() == dict(sorted(oc.dudt_nodes.items()))
    assert gc.get_neumann_edges() == oc.neumann_edges
    assert len(gc.functions) == len(oc.functions) == 5
    for i in range(pair.gtri.num_points):
        assert gc.has_condition(i) == oc.has_condition(i)
        if oc.is_dirichlet_node(i):
            assert gc.node_kind[i] == 1 and gc.node_fidx[i] == oc.dirichlet_nodes[i]
        elif oc.is_dudt_node(i):
            assert gc.node_kind[i] == 2 and gc.node_fidx[i] == oc.dudt_nodes[i]
    assert gc.has_constrained_edges() and gc.has_neumann_edges() and gc.has_dudt_nodes()


def test_constructor_errors_mirror_the_reference():
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 5, 5, single_boundary=True))
    mesh = pair.gmesh
    BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    with pytest.raises(AssertionError):  # problem.jl:131-132
        G.FVMProblem(mesh, BCs, diffusion_function=1.0, initial_condition=np.zeros(24), final_time=1.0)
    with pytest.raises(AssertionError):  # one function per boundary section
        G.BoundaryConditions(mesh, (G.Const(0.0), G.Const(1.0)), (G.Dirichlet, G.Dirichlet))
    with pytest.raises(G.UnsupportedClosureError):  # closures cannot run on the device
        G.FVMProblem(mesh, BCs, diffusion_function=lambda x, y, t, u, p: 1.0, initial_condition=np.zeros(25), final_time=1.0)
    with pytest.raises(G.UnsupportedClosureError):
        G.FVMProblem(mesh, BCs, diffusion_function=1.0, source_function=lambda x, y, t, u, p: u,
                     initial_condition=np.zeros(25), final_time=1.0)
    p1 = G.FVMProblem(mesh, BCs, diffusion_function=1.0, initial_condition=np.zeros(25), final_time=1.0)
    p2 = G.FVMProblem(mesh, BCs, diffusion_function=1.0, initial_condition=np.zeros(25), final_time=2.0)
    with pytest.raises(AssertionError):  # problem.jl:249-271
        G.FVMSystem(p1, p2)
    with pytest.raises(AssertionError):
        G.FVMSystem()
    with pytest.raises(G.InvalidFluxError, match="flux function q"):  # test/equations.jl:108: FVMSystem(prob, prob)
        G.FVMSystem(p1, p1)
    q1 = G.FVMProblem(mesh, BCs, flux_function=G.ConstantDiffusion(1.0), initial_condition=np.zeros(25), final_time=1.0)
    sys_ = G.FVMSystem(q1, q1)
    assert sys_.initial_condition.shape == (25, 2) and sys_.cnum_fncs == (0, 1)
    assert "FVMProblem with 25 nodes" in repr(p1)


def test_show_strings_mirror_the_reference():
    """The text/plain `show` methods the reference's own tests compare verbatim: test/conditions.jl:39-46 (BoundaryConditions,
    InternalConditions, Conditions), test/problem.jl:18-19,80-81,97-98 (FVMProblem, SteadyFVMProblem, FVMSystem), and
    geometry.jl:50-55 (FVMGeometry: solid vertices, triangles, edges)."""
    tri = G.triangulate_rectangle(0, 1, 0, 1, 5, 4, single_boundary=False)
    mesh = G.FVMGeometry(tri)
    assert repr(mesh) == "FVMGeometry with 20 control volumes, 24 triangles, and 43 edges"  # E = N + T - 1
    types = (G.Neumann, G.Dirichlet, G.Dudt, G.Constrained)
    BCs = G.BoundaryConditions(mesh, (G.Const(0.0),) * 4, types)
    assert repr(BCs) == "BoundaryConditions with 4 boundary conditions with types (Neumann, Dirichlet, Dudt, Constrained)"
    one = G.FVMGeometry(G.triangulate_rectangle(0, 1, 0, 1, 5, 4, single_boundary=True))
    assert repr(G.BoundaryConditions(one, G.Const(0.0), G.Dirichlet)) == "BoundaryConditions with 1 boundary condition with type Dirichlet"
    ICs = G.InternalConditions((G.Const(1.0),), dirichlet_nodes={6: 0, 7: 0}, dudt_nodes={11: 0})
    assert repr(ICs) == "InternalConditions with 2 Dirichlet nodes and 1 Dudt nodes"
    conds = G.Conditions(mesh, BCs, ICs)
    # bottom: 4 Neumann edges; right: Dirichlet on its 4 nodes; top: Dudt on its 5 nodes, one of which (the top right
    # corner) is also a Dirichlet node and is counted in both Dicts like merge_conditions! does; left: 3 Constrained edges
    assert repr(conds) == "Conditions with\n   4 Neumann edges\n   3 Constrained edges\n   6 Dirichlet nodes\n   6 Dudt nodes"
    prob = G.FVMProblem(one, G.BoundaryConditions(one, G.Const(0.0), G.Dirichlet), diffusion_function=1.0, initial_condition=np.zeros(20),
                        initial_time=0.5, final_time=2.0)
    assert repr(prob) == "FVMProblem with 20 nodes and time span (0.5, 2.0)"
    assert repr(G.SteadyFVMProblem(prob)) == "SteadyFVMProblem with 20 nodes"
    q = G.FVMProblem(one, G.BoundaryConditions(one, G.Const(0.0), G.Dirichlet), flux_function=G.ConstantDiffusion(1.0), initial_condition=np.zeros(20),
                     final_time=2.0)
    system = G.FVMSystem(q, q, q)
    assert repr(system) == "FVMSystem with 3 equations and time span (0.0, 2.0)"
    assert repr(G.SteadyFVMProblem(system)) == "SteadyFVMProblem with 20 nodes and 3 equations"


def test_template_argument_errors_mirror_the_reference():
    """poissons_equation.jl:69-70, mean_exit_time.jl:66-69: ArgumentError before any device work."""
    pair = Pair(G.triangulate_rectangle(0, 1, 0, 1, 5, 5, single_boundary=False))
    mesh = pair.gmesh
    dudt = G.BoundaryConditions(mesh, (G.Const(0.0),) * 4, (G.Dirichlet, G.Dudt, G.Dirichlet, G.Dirichlet))
    cons = G.BoundaryConditions(mesh, (G.Const(0.0),) * 4, (G.Dirichlet, G.Constrained, G.Dirichlet, G.Dirichlet))
    with pytest.raises(ValueError, match="PoissonsEquation does not support Dudt nodes"):
        G.PoissonsEquation(mesh, dudt, source_function=lambda x, y, p: 0 * x)
    with pytest.raises(ValueError, match="PoissonsEquation does not support Dudt nodes"):  # sic, laplaces_equation.jl:62
        G.LaplacesEquation(mesh, dudt)
    with pytest.raises(ValueError, match="MeanExitTimeProblem does not support Dudt nodes"):
        G.MeanExitTimeProblem(mesh, dudt, diffusion_function=1.0)
    with pytest.raises(ValueError, match="MeanExitTimeProblem does not support Constrained edges"):
        G.MeanExitTimeProblem(mesh, cons, diffusion_function=1.0)
    with pytest.raises(ValueError):
        G.Tsit5(adaptive=False)
    assert G.Tsit5().adaptive and not G.Tsit5(0.1).adaptive


@pytest.mark.parametrize("tile", [64, 256, 1024])
def test_finalize_host_planning_invariants(tile):
    """fvm_plan_selftest runs the host half of fvm_finalize (Hilbert tiling, tile-major renumbering, tile-local
    ids, gather lists, interface / partial-slot bookkeeping, live boundary edges) WITHOUT a device and checks, in
    the library, every invariant the kernels rely on; here its summary is checked against the mesh."""
    import ctypes as C
    from fvm_b200 import _lib as L
    lib = L.lib()
    cases = []
    lat = G.triangulate_rectangle(0, 2, 0, 2, 50, 50, single_boundary=True)
    kinds = np.zeros(2500, np.uint8)
    kinds[np.unique(lat.boundary_edges()[0])] = 1  # README: all-Dirichlet boundary -> no live edge
    cases.append((lat, 1, kinds, 0))
    cases.append((lat, 2, None, 196))  # all-free boundary, 2 species: every boundary edge is live
    un = delaunay_mesh(1500, 3, extra_points=5)
    cases.append((un, 1, None, len(un.boundary_edges()[0])))
    big = G.triangulate_rectangle(0, 1, 0, 3, 120, 333, single_boundary=False)
    cases.append((big, 1, None, len(big.boundary_edges()[0])))
    for tri, neq, kind, n_live in cases:
        uv = L.i32(tri.boundary_edges()[0] + 1)  # 1-based, like the Julia side passes them
        pts, t1 = L.f64(tri.points), L.i32(tri.triangles + 1)
        kk = None if kind is None else np.ascontiguousarray(np.tile(kind, neq))
        st = np.zeros(8, np.int64)
        rc = lib.fvm_plan_selftest(L.dp(pts), tri.num_points, L.ip(t1), tri.num_triangles, 1, neq, L.ip(uv), len(uv), L.bp(kk), tile,
                                   st.ctypes.data_as(L.c_lp))
        assert rc == L.OK, lib.fvm_last_error(None).decode()
        n_tiles, n_vert, n_ifc, n_partial, n_ext, max_nloc, live, gather = st.tolist()
        assert n_tiles == -(-tri.num_triangles // tile) and gather == 3 * tri.num_triangles
        assert n_vert == int(tri.solid_vertex_mask().sum()) and live == n_live
        assert n_partial >= n_ifc + n_ext and max_nloc <= 3 * tile  # every interface node: own slot + one per foreign tile
        if n_tiles == 1:
            assert n_ext == 0 and n_ifc == len(np.unique(tri.boundary_edges()[0])) * (live > 0)
    # argument errors
    st = np.zeros(8, np.int64)
    assert lib.fvm_plan_selftest(L.dp(pts), tri.num_points, L.ip(t1), tri.num_triangles, 1, 1, L.ip(uv), len(uv), None, 100,
                                 st.ctypes.data_as(L.c_lp)) == L.ERR_ARG
    bad_uv = L.i32(uv[:, ::-1].copy())  # clockwise edges are not edges of any triangle
    assert lib.fvm_plan_selftest(L.dp(pts), tri.num_points, L.ip(t1), tri.num_triangles, 1, 1, L.ip(bad_uv), len(bad_uv), None, tile,
                                 st.ctypes.data_as(L.c_lp)) == L.ERR_ARG
    assert b"not a ccw edge" in lib.fvm_last_error(None)


def test_header_is_plain_c_and_the_c_example_runs(tmp_path):
    """include/fvmcuda.h must compile as strict C99 (it is what a Julia ccall / cgo / JNI binding reads), and
    examples/readme_diffusion.c drives the ABI from plain C: the host-only entry points work anywhere, the compute
    part ends with status 2 and the "no CPU fallback" message when there is no device (0 on a GPU box)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "readme_diffusion")
    libdir = os.path.dirname(G.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "readme_diffusion.c"), "-L", libdir, "-lfvmcuda", "-Wl,-rpath," + libdir, "-lm",
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert "plan: 19 tiles, 2500 vertices" in r.stdout and "0 live boundary edges" in r.stdout
    assert "partition: 625 625 625 625 nodes" in r.stdout and "written and verified" in r.stdout
    if _have_gpu():
        assert r.returncode == 0 and "Tsit5 to t = 0.5" in r.stdout, r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
    assert not os.path.exists(str(tmp_path / "readme_mesh.fvmw")) or r.returncode != 2


def test_sass_is_sm100a_without_fp64_atomics_and_with_bulk_copies():
    """What the design claims about the generated code, checked on the objects the library is linked from: every cubin is
    sm_100a; no kernel on the path contains an atomic or a reduction (the scatter is a fixed-order gather) -- the only
    ones in the library are the integer counters of the one-off incidence-list build (fvm_linear) and the halo epoch
    counters (fvm_shard), never fp64; the streaming RHS kernel moves its tile packs with bulk async copies completing on
    mbarriers (UBLKCP / SYNCS) and gathers external values with cp.async (LDGSTS)."""
    import re
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    from fvm_b200 import _lib as L
    L.lib()  # builds the objects if they are missing
    libdir = os.path.join(ROOT, "finitevolumemethod.jl_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "fvm_rhs_stream.o")):
        pytest.skip("object files are not in the tree (library built elsewhere)")
    atom = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?(ATOM|ATOMG|ATOMS|RED)\b(\S*)", re.M)
    counts = {}
    for name in sorted(f for f in os.listdir(libdir) if f.endswith(".o")):
        sass = subprocess.run([exe, "-sass", os.path.join(libdir, name)], capture_output=True, text=True, check=True).stdout
        archs = set(re.findall(r"arch = (sm_\w+)", sass))
        assert archs <= {"sm_100a"}, (name, archs)
        hits = atom.findall(sass)
        assert not any("F64" in suffix or "64.F" in suffix for _, suffix in hits), (name, hits)
        counts[name] = len(hits)
        if name == "fvm_rhs_stream.o":
            assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass and "LDGSTS" in sass
    assert {k for k, v in counts.items() if v} <= {"fvm_linear.o", "fvm_shard.o"}, counts
    for name in ("fvm_rhs.o", "fvm_rhs_stream.o", "fvm_solvers.o", "fvm_jacobian.o", "fvm_pipe.o"):
        assert counts[name] == 0
