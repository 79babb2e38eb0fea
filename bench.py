#!/usr/bin/env python
"""bench.py -- fvm_eqs! triangle-updates/s (+ template SpMV GB/s) on a 16.7M-node lattice, B200.

    python bench.py --gpus N --steps K --warmup W            # this framework (libfvmcuda)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on host cores

A step = one pass of the hot path over the whole mesh: one `fvm_eqs!` evaluation (boundary-edge,
tile and interface kernels) on the README diffusion problem scaled to BASELINE configs[1]'s mesh
(triangulate_rectangle 4096x4096 on [0,2]^2, Dirichlet u=0, D=1/9).  The headline variant is the
path the engine takes by default (geometry_mode 1: the streaming kernel that recomputes the geometry
from vertex coordinates, 12*T + 41*N algorithmic bytes); the stored-geometry layouts of north_star (a)
(reduced 108*T + 25*N, general 180*T + 25*N) are reported under "variants", every variant also as a
fraction of the roofline on the MINIMAL byte count (12*T + 41*N).  `value` times it with `u`
resident in HBM (native order); `e2e` times the public `fvm_eqs(du,u,p,t)` call with pinned HOST
buffers (H2D + permutation + kernels + D2H).  N>1: weak scaling, each rank owns a 4096-row strip, plus
`strong_8192` = BASELINE config 5 (the 8192x8192 lattice split over the N GPUs: RHS, template SpMV and
Tsit5 step) and `parity_max_rel` (owned rows at the cuts against a single-domain evaluation).
One JSON line on stdout (rank 0)."""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240517


def profiled_traffic(key):
    """DRAM bytes per launch of the dominant kernel FROM THE COMMITTED ncu CAPTURE of the same command
    (profiles/r02_traffic.json, falling back to round 1's) -- evidence of that capture, not of this run."""
    for f in ("r02_traffic.json", "r01_traffic.json"):
        try:
            return int(json.load(open(os.path.join(ROOT, "profiles", f)))[key]["traffic_bytes"]), "profiles/" + f
        except Exception:
            continue
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def lattice_problem(G, nx, ny, flux, rank=0, world=1):
    """world == 1: the [0,2]^2 lattice.  world > 1 (weak scaling): the rank-local strip (owned rows +
    one ghost row per neighbour) of the nx x (ny*world) lattice on [0,2] x [0,2*world]."""
    local = None
    if world == 1:
        tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, ny, single_boundary=True)
    else:
        local = G.lattice_strip_local(0.0, 2.0, 0.0, 2.0 * world, nx, ny, rank, world)
        tri = local.triangulation
    mesh = G.FVMGeometry(tri)
    BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    prob = G.FVMProblem(mesh, BCs, diffusion_function=flux, initial_condition=ic, final_time=0.5)
    return prob, local


def rhs_bytes(T, N, neq, layout):
    """Algorithmic bytes of one RHS (BASELINE.md section 2 / SURVEY.md 8d)."""
    if layout == "general":      # north_star SoA: 3 idx + 9 s + 6 midpoints + 6 scaled normals
        return 180 * T + (17 * neq + 8) * N
    if layout == "reduced":      # flux independent of (x,y,u): s1..s6 + scaled normals
        return 108 * T + (17 * neq + 8) * N
    if layout == "recompute":    # geometry recomputed from vertex coordinates
        return 12 * T + (24 + 17 * neq) * N
    raise ValueError(layout)


VARIANTS = {
    # name: (flux spec factory, geometry_mode, byte layout)
    "general_stored": (lambda G: G.PowerDiffusion(1 / 9, 1.0), 0, "general"),
    "const_stored": (lambda G: G.ConstantDiffusion(1 / 9), 0, "reduced"),
    "general_recompute": (lambda G: G.PowerDiffusion(1 / 9, 1.0), 1, "recompute"),
    "const_recompute": (lambda G: G.ConstantDiffusion(1 / 9), 1, "recompute"),
}


def time_rhs(torch, eng, u_d, du_d, steps, warmup):
    """CUDA events on the handle's stream; returns (ms per step, dominant-kernel ms per launch)."""
    stream = torch.cuda.ExternalStream(eng.stream())
    for _ in range(warmup):
        eng.rhs_device(du_d.data_ptr(), u_d.data_ptr(), 0.0, native=True)
    eng.synchronize()
    eng.set_profiling(2 * steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.rhs_device(du_d.data_ptr(), u_d.data_ptr(), 0.0, native=True)
    e1.record(stream)
    eng.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms, kn = eng.get_profile()
    eng.set_profiling(0)
    return ms, (kms / steps if kn else float("nan"))  # per step (the overlapped sharded schedule launches it twice)


def time_template(torch, G, prob, nx, steps, warmup, peak, lmesh=None, dist=None, e2e=True):
    """BASELINE configs[1]: DiffusionEquation template on the same mesh: y = A x + b SpMV and the
    device-resident fixed-step Tsit5 (6 SpMV + 6 stage combinations per step)."""
    mesh = prob.mesh
    tri = mesh.triangulation
    BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    t0 = time.perf_counter()
    tpl = G.DiffusionEquation(mesh, BCs, diffusion_function=1 / 9, initial_condition=prob.initial_condition, final_time=1.0,
                              ghost=None if lmesh is None else lmesh.is_ghost, tile_triangles=int(os.environ.get("FVM_TPL_TILE", "0")))
    if lmesh is not None:
        G.install_halo(tpl.engine, lmesh, dist)
    setup_s = time.perf_counter() - t0
    eng = tpl.engine
    N = eng.N
    st = eng.stats()
    nnz = st["nnz"]
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED)
    x = 50.0 * torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty_like(x)
    stream = torch.cuda.ExternalStream(eng.stream())
    for _ in range(warmup):
        eng.spmv_device(y.data_ptr(), x.data_ptr(), True, native=True)
    eng.synchronize()
    eng.set_profiling(2 * steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    e0.record(stream)
    for _ in range(steps):
        eng.spmv_device(y.data_ptr(), x.data_ptr(), True, native=True)
    e1.record(stream)
    eng.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    kms, kn = eng.get_profile()
    eng.set_profiling(0)
    kms = kms / steps if kn else float("nan")
    B = 12 * nnz + 4 * (N + 1) + 24 * N
    # fixed-step Tsit5 at a stable dt (|lambda|max ~ 8 D / h^2, dt < 3.3 / |lambda|max)
    h = 2.0 / (nx - 1)
    dt = 0.2 * 3.3 * h * h / (8.0 / 9.0)
    nst = 20
    u = torch.from_numpy(tpl.u0).cuda()
    eng.tsit5_device(u.data_ptr(), 0.0, 2 * dt, dt, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.tsit5_device(u.data_ptr(), 0.0, nst * dt, dt, True)
    torch.cuda.synchronize()
    ts_ms = (time.perf_counter() - t0) / nst * 1e3
    # the operator through the public call with pinned HOST vectors (mul!(du, A, u) of a host integrator)
    e2e_spmv_ms = None
    if dist is None and e2e:
        xh = torch.empty(N, dtype=torch.float64).pin_memory()
        yh = torch.empty(N, dtype=torch.float64).pin_memory()
        xh.copy_(x.cpu())
        tpl.mul(yh.numpy(), xh.numpy())
        t0 = time.perf_counter()
        for _ in range(5):
            tpl.mul(yh.numpy(), xh.numpy())
        e2e_spmv_ms = (time.perf_counter() - t0) / 5 * 1e3
        del xh, yh
    # the SpMV is ONE launch (tile CTAs on interior rows + a few CTAs on the sliced-ELL tail of interface rows)
    out = {"spmv_e2e_ms": e2e_spmv_ms,"spmv_ms": ms, "spmv_kernel_ms": kms, "spmv_gbs": B / ms / 1e6, "spmv_frac": B / ms / 1e6 / peak, "spmv_alg_bytes": B,
           "nnz": nnz, "assemble_setup_s": setup_s, "tsit5_ms_per_step": ts_ms, "tsit5_steps": nst, "tsit5_dt": dt,
           "tsit5_launches": nst * 12 + 3, "finite": bool(torch.isfinite(u).all().item())}
    eng.close()
    return out


def time_extras(torch, G, nx, steps, warmup, peak):
    """The other BASELINE configs, at full size on one GPU: config 4 (2-species Keller-Segel FVMSystem
    RHS, 4096^2), config 3 (MeanExitTimeProblem, 2048^2, Jacobi-PCG) and config 1 (README 50x50)."""
    out = {}
    # ---- config 4: FVMSystem, registered u-dependent flux, all-Neumann zero flux -------------------
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    N = tri.num_points
    rng = np.random.default_rng(SEED)
    ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann)
    pu = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=0.01 * rng.random(N), final_time=1.0)
    pv = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=np.zeros(N), final_time=1.0)
    p = G.get_cuda_parameters(G.FVMSystem(pu, pv))
    eng = p.engine
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED)
    u_d = 0.01 * torch.rand(2 * N, dtype=torch.float64, device="cuda", generator=g)
    du_d = torch.empty_like(u_d)
    ms, kms = time_rhs(torch, eng, u_d, du_d, steps, warmup)
    B = rhs_bytes(eng.T, N, 2, "recompute")  # the engine's default path: geometry recomputed, 12*T + (24 + 17*2)*N
    out["system_rhs"] = {"config": "FVMSystem 2-species Keller-Segel (chi(u) grad v - grad u; -D grad v), all-Neumann, %dx%d" % (nx, nx),
                         "mtri_s": eng.T / ms / 1e3, "ms_per_step": ms, "tile_kernel_ms": kms, "alg_bytes": B, "kernel": "rhs_stream_kernel",
                         "bytes_formula": "12*T + 58*N", "gbs": B / kms / 1e6, "frac": B / kms / 1e6 / peak,
                         "stored_layout_equivalent_gbs": (180 * eng.T + 42 * N) / kms / 1e6,
                         "finite": bool(torch.isfinite(du_d).all().item())}
    eng.close()
    del u_d, du_d, p, eng
    # ---- config 3: steady MeanExitTimeProblem, 2048^2, Jacobi-PCG -----------------------------------
    n3 = max(64, nx // 2)
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, n3, n3, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    t0 = time.perf_counter()
    met = G.MeanExitTimeProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    sol = G.solve(met, G.KrylovJacobi("pcg", rtol=1e-10, maxiter=40000))
    t_solve = time.perf_counter() - t0
    # closed form on the square [0,L]^2: T(x,y) = (16 L^2 / (D pi^4)) sum_odd sin(m pi x/L) sin(n pi y/L) / (m n (m^2+n^2))
    c = n3 // 2 * n3 + n3 // 2
    L = 2.0
    mm = np.arange(1, 400, 2, dtype=np.float64)
    xc, yc = tri.points[c]
    series = 16 * L * L / ((1 / 9) * np.pi**4) * np.sum(
        np.sin(mm[:, None] * np.pi * xc / L) * np.sin(mm[None, :] * np.pi * yc / L) / (mm[:, None] * mm[None, :] * (mm[:, None]**2 + mm[None, :]**2)))
    out["steady_pcg"] = {"config": "MeanExitTimeProblem %dx%d, D=1/9, Jacobi-PCG rtol 1e-10" % (n3, n3), "iters": sol.iters,
                         "relres": sol.relres, "solve_s": t_solve, "assemble_s": t_asm, "ms_per_iter": 1e3 * t_solve / max(1, sol.iters),
                         "centre_value": float(sol.u[c]), "centre_closed_form": float(series)}
    met.engine.close()
    # ---- config 1: README, 50x50, Tsit5 to t = 0.5 (launch-bound) -----------------------------------
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, 50, 50, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    prob = G.FVMProblem(mesh, BCs, diffusion_function=G.ConstantDiffusion(1 / 9), initial_condition=ic, final_time=0.5)
    pp = G.get_cuda_parameters(prob)
    G.solve(prob, G.Tsit5(0.0025), p=pp)
    t0 = time.perf_counter()
    G.solve(prob, G.Tsit5(0.0025), p=pp)
    t_fvm = time.perf_counter() - t0
    tpl = G.DiffusionEquation(mesh, BCs, diffusion_function=1 / 9, initial_condition=ic, final_time=0.5)
    G.solve(tpl, G.Tsit5(0.0025))
    t0 = time.perf_counter()
    G.solve(tpl, G.Tsit5(0.0025))
    t_tpl = time.perf_counter() - t0
    out["readme_50x50"] = {"config": "README diffusion, Tsit5 fixed dt=0.0025 to t=0.5 (200 steps)", "fvmproblem_solve_ms": 1e3 * t_fvm,
                           "template_solve_ms": 1e3 * t_tpl}
    pp.engine.close()
    tpl.engine.close()
    return out


def host_threads():
    """Threads the CPU arm uses: every core the process may run on.  torchrun exports OMP_NUM_THREADS=1 to its
    workers; a baseline on one core would be meaningless, so the count is set explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def julia_probe():
    """SURVEY 8c / BASELINE.md 3.1: use the real reference if Julia exists on the bench host.  It never has so far."""
    import shutil
    exe = shutil.which("julia")
    if not exe:
        return {"found": False, "note": "`command -v julia` found nothing on this host: the CPU arm is the C port of the reference's structure (kind: port)"}
    return {"found": True, "path": exe, "note": "julia is on PATH but FiniteVolumeMethod.jl's dependencies (DelaunayTriangulation, SciMLBase, ...) "
                                                "are not vendored and there is no network: the CPU arm stays the C port"}


def cpu_baseline(nx, steps=None, warmup=1, budget_s=25.0, extras=True):
    """The oracle port (oracle/fvm_oracle_c.c, reference-structured: hash-table triangle props, per-thread du
    copies, serial combine) timed on ALL host cores on the README problem of side nx; beside it the flat-array
    variant, the CPU CSR SpMV (1 thread like SparseArrays' mul!, and all threads) and a SuperLU steady solve."""
    from oracle.c_oracle import COracle
    import fvm_b200 as G
    nthreads = host_threads()
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True)
    uv, _ = tri.boundary_edges()
    co = COracle(tri.points, tri.triangles, np.unique(uv), 1 / 9, nthreads=nthreads)
    u = 50 * np.random.default_rng(SEED).random(tri.num_points)
    du = np.empty_like(u)
    for _ in range(max(1, warmup)):
        co.fvm_eqs_threaded(u, du)
    t0 = time.perf_counter()
    co.fvm_eqs_threaded(u, du)
    one = time.perf_counter() - t0
    reps = steps if steps else max(3, min(200, int(budget_s / max(one, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(reps):
        co.fvm_eqs_threaded(u, du)
    dt = (time.perf_counter() - t0) / reps
    T = tri.num_triangles
    out = {"value": T / dt / 1e6, "unit": "Mtriangle-updates/s", "cores": nthreads, "kind": "port",
           "sample": "README diffusion fvm_eqs! on the %dx%d lattice (%d triangles), %d threaded calls of the reference-structured "
                     "C port (Dict-style triangle props, per-thread du copies, serial combine)" % (nx, nx, T, reps),
           "ms_per_step": dt * 1e3, "nx": nx, "julia": julia_probe()}
    if extras:
        nflat = max(2, min(reps, 10))
        co.fvm_eqs_flat(u, du)
        t0 = time.perf_counter()
        for _ in range(nflat):
            co.fvm_eqs_flat(u, du)
        out["cpu_flat"] = {"value": T / ((time.perf_counter() - t0) / nflat) / 1e6, "unit": "Mtriangle-updates/s", "cores": nthreads,
                           "note": "same arithmetic on flat arrays without the hash containers, parallel combine (BASELINE.md 3: fair CPU line)"}
        try:  # the template SpMV on the CPU: 7-point operator of the same lattice in CSR
            from oracle.c_oracle import spmv as c_spmv
            import scipy.sparse as sp
            N = tri.num_points
            A = sp.diags([np.full(N, -4.0), np.ones(N - 1), np.ones(N - 1), np.ones(N - nx), np.ones(N - nx), np.full(N - nx + 1, 1e-16),
                          np.full(N - nx + 1, 1e-16)], [0, 1, -1, nx, -nx, nx - 1, -(nx - 1)], format="csr")
            rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
            bb = np.zeros(N)
            Bs = 12 * A.nnz + 4 * (N + 1) + 24 * N
            for nt, key in ((1, "cpu_spmv_1t"), (nthreads, "cpu_spmv_nt")):
                c_spmv(rp, ci, va, bb, u, nt)
                t0 = time.perf_counter()
                for _ in range(5):
                    c_spmv(rp, ci, va, bb, u, nt)
                ts = (time.perf_counter() - t0) / 5
                out[key] = {"gbs": Bs / ts / 1e9, "ms": ts * 1e3, "threads": nt,
                            "note": "CSR y = A x + b, %s" % ("single-threaded like SparseArrays' mul! (diffusion_equation.jl:93-94)" if nt == 1 else "all host threads")}
            del A, rp, ci, va
        except Exception as e:  # the RHS line is the baseline; the SpMV lines are extra
            out["cpu_spmv_note"] = "unavailable (%s)" % type(e).__name__
    co.close()
    return out


def cpu_steady_baseline(n=512):
    """BASELINE config 3 beside the GPU Jacobi-PCG: sparse-direct solve (SciPy SuperLU standing in for the reference's
    KLUFactorization, docs/src/literate_wyos/poissons_equation.jl:101) of the MeanExitTimeProblem on an n x n lattice,
    assembled by the oracle (vectorised restatement of abstract_templates.jl:73-99)."""
    from oracle import fvm_oracle as O
    tri = O.triangulate_rectangle(0, 2, 0, 2, n, n, single_boundary=True)
    mesh = O.FVMGeometry(tri)
    BCs = O.BoundaryConditions(mesh, (lambda x, y, t, u, p: 0.0,), (O.Dirichlet,))
    ref = O.MeanExitTimeProblem(mesh, BCs, diffusion_function=lambda x, y, p: 1 / 9, vectorised=True)
    t0 = time.perf_counter()
    x = O.solve_steady(ref)
    return {"n": n, "unknowns": n * n, "superlu_s": time.perf_counter() - t0, "centre_value": float(x[(n // 2) * n + n // 2]), "threads": 1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the CPU arm runs the SAME mesh as the GPU arm when K steps of it fit in a few minutes (~1 s per 4096^2 call)
    nx = args.ref_nx if args.ref_nx else (args.nx if (args.steps + args.warmup) <= 120 else max(1024, args.nx // 2))
    cb = cpu_baseline(nx, steps=args.steps, warmup=args.warmup, extras=False)
    same = nx == args.nx
    line = {"impl": "reference", "metric": "fvm_eqs! Mtriangle-updates/s", "value": cb["value"], "unit": cb["unit"],
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "README diffusion fvm_eqs! (D=1/9, Dirichlet u=0) on triangulate_rectangle %dx%d; CPU arm: C port of the "
                                   "reference's threaded path on %d host threads%s" % (nx, nx, cb["cores"], "" if same else
                                   " (bounded sample of the %dx%d lattice: the metric is per triangle)" % (args.nx, args.nx)),
                       "same_mesh_as_gpu_arm": same, "julia": cb["julia"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def global_u(gids):
    """u as a function of the GLOBAL node id (any rank can evaluate any node): 50 * frac(sin(g) * 43758.5453)."""
    g = np.asarray(gids, dtype=np.float64)
    v = np.sin(g * 12.9898 + 78.233) * 43758.5453
    return 50.0 * (v - np.floor(v))


def sharded_parity(torch, G, eng, lmesh, du_d, nx, ny_total, ymax, flux, rank, world, gmode):
    """In-run parity of a sharded RHS: du of the first, a middle and the last OWNED row of this rank's strip (the rows
    that read the ghost layer) against a single-domain evaluation of a 5-row patch of the global lattice on this GPU
    (same coordinates, same u by global id).  Returns max relative difference (inf-norm per row)."""
    owned_rows = lmesh.n_owned // nx
    j0 = rank * owned_rows
    du_caller = torch.empty_like(du_d)
    from fvm_b200 import _lib as L
    L.check(eng.h, L.lib().fvm_from_native(eng.h, du_d.data_ptr(), du_caller.data_ptr()))
    eng.synchronize()
    du_h = du_caller.cpu().numpy()
    worst = 0.0
    for j in sorted({j0, j0 + owned_rows // 2, j0 + owned_rows - 1}):
        r0, r1 = max(0, j - 2), min(ny_total, j + 3)
        tri = G.lattice_rows(0.0, 2.0, 0.0, ymax, nx, ny_total, r0, r1)
        mesh = G.FVMGeometry(tri)
        gid = (np.arange(r0, r1, dtype=np.int64)[:, None] * nx + np.arange(nx)[None, :]).ravel()
        up = global_u(gid)
        prob = G.FVMProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=flux,
                            initial_condition=up, final_time=0.5)
        pp = G.get_cuda_parameters(prob, geometry_mode=gmode)
        ref = G.fvm_eqs(np.empty_like(up), up, pp, 0.0)[(j - r0) * nx:(j - r0 + 1) * nx]
        pp.engine.close()
        got = du_h[(j - j0) * nx:(j - j0 + 1) * nx]  # owned rows come first, in ascending global order
        if j in (0, ny_total - 1):
            continue  # a global boundary row is Dirichlet in both: du = 0
        worst = max(worst, float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)))
    return worst


def patch_parity(torch, G, eng, tri_g, lmesh, du_d, flux, gmode, rank, max_cut=3000, max_inner=1000):
    """In-run parity of a run sharded by an ARBITRARY partition: du at this rank's owned nodes next to the cuts (the nodes
    whose triangle fans read the exchanged ghost values; a sample of at most `max_cut`) and at random owned nodes, against a
    single-domain evaluation on this GPU of the patch made of every triangle incident to those nodes (global coordinates,
    same u by global id; patch rim = Dirichlet, ignored).  Returns max |difference| / max |reference| over the sample."""
    from fvm_b200 import _lib as L
    rng = np.random.default_rng(SEED + rank)
    cut = np.unique(np.concatenate([np.asarray(x) for x in lmesh.send_nodes])) if lmesh.send_nodes else np.zeros(0, np.int64)
    if len(cut) > max_cut:
        cut = rng.choice(cut, max_cut, replace=False)
    inner = rng.choice(lmesh.n_owned, min(max_inner, lmesh.n_owned), replace=False)
    loc = np.unique(np.concatenate([cut, inner]).astype(np.int64))          # local ids, all owned
    patch, verts, pl = G.patch_mesh(tri_g, lmesh.global_nodes[loc])
    up = global_u(verts)
    mesh = G.FVMGeometry(patch)
    prob = G.FVMProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=flux, initial_condition=up, final_time=0.5)
    pp = G.get_cuda_parameters(prob, geometry_mode=gmode)
    ref = G.fvm_eqs(np.empty_like(up), up, pp, 0.0)[pl]
    pp.engine.close()
    du_caller = torch.empty_like(du_d)
    L.check(eng.h, L.lib().fvm_from_native(eng.h, du_d.data_ptr(), du_caller.data_ptr()))
    eng.synchronize()
    got = du_caller[torch.from_numpy(loc).cuda()].cpu().numpy()
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)), len(loc)


def strong_scaling_leg(torch, G, dist, rank, world, local, steps, warmup, peak, nxs=8192, partition="strips"):
    """BASELINE config 5 as written: the nxs x nxs lattice (67M nodes, 134M triangles at 8192) on [0,2]^2 split over the
    `world` GPUs, README diffusion: fvm_eqs!, the DiffusionEquation template SpMV and a Tsit5 step, device-timed, max over
    ranks.  partition: "strips" (row strips, built without the global mesh), "rcb" (recursive coordinate bisection) or
    "graph" (fvm_partition_graph, the METIS-style multilevel partitioner; every rank computes the same deterministic
    partition of the global mesh and extracts its part).  world == 1 gives the 1-GPU leg the speed-ups are quoted against."""
    t0 = time.perf_counter()
    tri_g, part_s, cut_edges = None, 0.0, None
    if partition == "strips" or world == 1:
        lmesh = G.lattice_strip_local(0.0, 2.0, 0.0, 2.0, nxs, nxs // world, rank, world)
    else:
        tri_g = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nxs, nxs, single_boundary=True)
        tp = time.perf_counter()
        owner = G.partition_rcb(tri_g.points, world) if partition == "rcb" else G.partition_graph(tri_g, world)
        part_s = time.perf_counter() - tp
        if rank == 0:
            cut_edges = G.edge_cut(tri_g, owner)
        lmesh = G.extract_local(tri_g, owner, rank, world)
        del owner
    tri = lmesh.triangulation
    mesh = G.FVMGeometry(tri)
    BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
    u_h = global_u(lmesh.global_nodes)
    prob = G.FVMProblem(mesh, BCs, diffusion_function=G.ConstantDiffusion(1 / 9), initial_condition=u_h, final_time=0.5)
    if world == 1:
        p = G.get_cuda_parameters(prob, device=local)
    else:
        p = G.get_sharded_cuda_parameters(prob, lmesh, dist, device=local)
    eng = p.engine
    setup_rhs = time.perf_counter() - t0
    N = eng.N
    Tg = 2 * (nxs - 1) * (nxs - 1)
    u_c = torch.from_numpy(u_h).cuda()
    u_d = torch.empty_like(u_c)
    from fvm_b200 import _lib as L
    L.check(eng.h, L.lib().fvm_to_native(eng.h, u_c.data_ptr(), u_d.data_ptr()))
    du_d = torch.empty_like(u_d)
    if dist is not None:
        dist.barrier()
    ms_rhs, kms = time_rhs(torch, eng, u_d, du_d, steps, warmup)
    parity, parity_nodes = None, None
    if world > 1 and tri_g is None:
        parity = sharded_parity(torch, G, eng, lmesh, du_d, nxs, nxs, 2.0, G.ConstantDiffusion(1 / 9), rank, world, 1)
    elif world > 1:
        parity, parity_nodes = patch_parity(torch, G, eng, tri_g, lmesh, du_d, G.ConstantDiffusion(1 / 9), 1, rank)
    halo_nodes = int(sum(len(x) for x in lmesh.recv_nodes))
    n_neigh = len(lmesh.neighbours)
    eng.close()
    del p, eng, u_d, du_d, u_c, tri_g
    torch.cuda.empty_cache()
    tpl = time_template(torch, G, prob, nxs, steps, warmup, peak, lmesh if world > 1 else None, dist, e2e=False)
    out = {"partition": partition if world > 1 else "none", "rhs_ms": ms_rhs, "spmv_ms": tpl["spmv_ms"], "tsit5_ms_per_step": tpl["tsit5_ms_per_step"],
           "setup_s": setup_rhs, "partition_s": part_s, "assemble_setup_s": tpl["assemble_setup_s"], "nodes_per_gpu": N, "ghost_nodes_per_gpu": halo_nodes,
           "neighbours": n_neigh, "parity_max_rel": parity}
    if cut_edges is not None:
        out["edge_cut"] = int(cut_edges)
    if parity_nodes is not None:
        out["parity_nodes_per_rank"] = parity_nodes
    if dist is not None:
        v = torch.tensor([out["rhs_ms"], out["tsit5_ms_per_step"], parity if parity is not None else 0.0, float(halo_nodes), float(n_neigh)],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        out["rhs_ms"], out["tsit5_ms_per_step"], out["parity_max_rel"] = float(v[0]), float(v[1]), float(v[2])
        out["ghost_nodes_per_gpu"], out["neighbours"] = int(v[3]), int(v[4])  # max over ranks
    out["mesh"] = "%dx%d lattice on [0,2]^2 (%d nodes, %d triangles in total), %s" % (
        nxs, nxs, nxs * nxs, Tg, "one GPU" if world == 1 else
        {"strips": "%d row strips" % world, "rcb": "%d parts by recursive coordinate bisection" % world,
         "graph": "%d parts by fvm_partition_graph (multilevel: heavy-edge matching, graph growing, FM)" % world}[partition] + ", one-layer node halo")
    out["rhs_mtri_s"] = Tg / out["rhs_ms"] / 1e3
    nnz = nxs * nxs + 2 * (nxs * nxs + Tg - 1)
    out["spmv_gbs"] = (12 * nnz + 4 * (nxs * nxs + 1) + 24 * nxs * nxs) / out["spmv_ms"] / 1e6
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ref-nx", type=int, default=0, help="side of the CPU arm's lattice (0: the GPU arm's mesh when the run stays within minutes)")
    ap.add_argument("--no-strong", action="store_true", help="skip the BASELINE config 5 leg (8192^2 lattice over the N GPUs)")
    ap.add_argument("--strong-nx", type=int, default=8192)
    ap.add_argument("--strong-partition", default="strips", choices=["strips", "rcb", "graph"],
                    help="how the config-5 lattice is split over the GPUs (rcb / graph build the global mesh on every rank: "
                         "about a minute of extra host time at 8192^2)")
    ap.add_argument("--variant", default="const_recompute", choices=sorted(VARIANTS),
                    help="headline variant; const_recompute is what the engine runs by default for the README problem "
                         "(D = 1/9, a ConstantDiffusion, geometry_mode 1: streaming recompute kernel)")
    ap.add_argument("--all-variants", action="store_true")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-template", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the 400-launch pre-warm (for ncu captures)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # libraries (NCCL's version banner, torchrun) may write to fd 1: keep stdout for the ONE JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # one process per GPU on ONE host: give every rank its own slice of the cores (host-side planning in fvm_finalize is
    # OpenMP code and torchrun exports OMP_NUM_THREADS=1; pinned buffers are first-touched by the rank's own cores)
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            mine = cores[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            cores = mine
        os.environ["OMP_NUM_THREADS"] = str(len(cores))
    except Exception:
        pass
    import torch
    import fvm_b200 as G
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libfvmcuda has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["FVM_DEVICE"] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nx = args.nx
    peak, peak_src = measured_peak()
    # default: the README problem as the engine runs it (reduced stream) plus the same mesh through the
    # general 21-component layout of north_star (a); --all-variants adds the recompute-geometry kernels
    names = sorted(VARIANTS) if args.all_variants or world == 1 else [args.variant]
    if args.variant in names:
        names.remove(args.variant)
        names.append(args.variant)  # headline variant last: its engine stays alive for e2e
    results = {}
    eng = p = None
    parity_max_rel = None
    halo_mode = None
    for name in names:
        flux_f, gmode, layout = VARIANTS[name]
        if eng is not None:
            eng.close()
        t0 = time.perf_counter()
        prob, lmesh = lattice_problem(G, nx, nx, flux_f(G), rank, world)
        if lmesh is None:
            p = G.get_cuda_parameters(prob, tile_triangles=args.tile, geometry_mode=gmode, device=local)
        else:
            p = G.get_sharded_cuda_parameters(prob, lmesh, dist, tile_triangles=args.tile, geometry_mode=gmode, device=local)
        eng = p.engine
        setup_s = time.perf_counter() - t0
        N = eng.N
        # triangles per rank of the GLOBAL mesh (cut triangles, computed on both sides, count once)
        T = eng.T if world == 1 else 2 * (nx - 1) * (nx * world - 1) // world
        if lmesh is None:
            g = torch.Generator(device="cuda")
            g.manual_seed(SEED + rank)
            u_d = 50.0 * torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
        else:  # u is a function of the GLOBAL node id, so that any rank can evaluate any node (in-run parity check)
            from fvm_b200 import _lib as L
            u_c = torch.from_numpy(global_u(lmesh.global_nodes)).cuda()
            u_d = torch.empty_like(u_c)
            L.check(eng.h, L.lib().fvm_to_native(eng.h, u_c.data_ptr(), u_d.data_ptr()))
            eng.synchronize()
            del u_c
        du_d = torch.empty_like(u_d)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if rank == 0 and name == names[-1] else None
        if name == names[-1]:
            # nvidia-smi needs ~0.3 s to start sampling: keep the GPU under the same load meanwhile.  A FIXED
            # number of calls on EVERY rank: sharded calls contain NCCL send/recv and must pair up across ranks.
            for _ in range(0 if args.quick else 400):
                eng.rhs_device(du_d.data_ptr(), u_d.data_ptr(), 0.0, native=True)
            eng.synchronize()
        ms, kms = time_rhs(torch, eng, u_d, du_d, args.steps, args.warmup)
        torch.cuda.synchronize()
        if dist is not None:
            tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        clocks = sampler.stop() if sampler else None
        if lmesh is not None and name == names[-1]:
            halo_mode = eng.halo_mode()[0]
            parity = sharded_parity(torch, G, eng, lmesh, du_d, nx, nx * world, 2.0 * world, flux_f(G), rank, world, gmode)
            pv = torch.tensor([parity], dtype=torch.float64, device="cuda")
            dist.all_reduce(pv, op=dist.ReduceOp.MAX)
            parity_max_rel = float(pv.item())
        B = rhs_bytes(T, N, 1, layout)
        Bmin = rhs_bytes(T, N, 1, "recompute")
        results[name] = {"ms_per_step": ms, "mtri_s": world * T / ms / 1e3, "kernel_ms": kms, "layout": layout,
                         "alg_bytes": B, "kernel_gbs": B / kms / 1e6, "frac": B / kms / 1e6 / peak,
                         "frac_vs_minimal_bytes": Bmin / kms / 1e6 / peak, "step_frac_vs_minimal_bytes": Bmin / ms / 1e6 / peak,
                         "setup_s": setup_s, "tiles": eng.stats()}
        if rank == 0:
            sys.stderr.write("[bench] %-18s %.3f ms/step  %.1f Mtri/s  tile-kernel %.3f ms  %.0f GB/s (%.2f of %s)\n"
                             % (name, ms, results[name]["mtri_s"], kms, results[name]["kernel_gbs"], results[name]["frac"], peak_src))
    head = results[args.variant]

    # ---- e2e: the public call with pinned host buffers -------------------------------------
    u_h = torch.empty(N, dtype=torch.float64).pin_memory()
    du_h = torch.empty(N, dtype=torch.float64).pin_memory()
    u_h.copy_(u_d.cpu())
    un, dun = u_h.numpy(), du_h.numpy()
    for _ in range(3):  # warm-up, and the handle's one-off choice between the banded pipeline and the plain schedule
        G.fvm_eqs(dun, un, p, 0.0)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        G.fvm_eqs(dun, un, p, 0.0)  # synchronous: H2D, to-native, kernels, from-native, D2H
    e2e_ms = (time.perf_counter() - t0) / args.e2e_steps * 1e3
    # the same call without the banded copy/compute pipeline (fvm_pipe.cu), for reference; every rank makes the
    # same calls (sharded calls contain the NCCL halo exchange)
    os.environ["FVM_NO_PIPELINE"] = "1"
    G.fvm_eqs(dun, un, p, 0.0)
    t0 = time.perf_counter()
    for _ in range(3):
        G.fvm_eqs(dun, un, p, 0.0)
    e2e_plain_ms = (time.perf_counter() - t0) / 3 * 1e3
    del os.environ["FVM_NO_PIPELINE"]
    # the host link's own ceiling for this call: the same pinned buffers copied in and out on two streams with no kernel
    # at all, every rank at once (on the 8-GPU VM all ranks share one 154 GB/s PCIe root: profiles/r02_host_link_probe_8gpu.log)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    scratch = torch.empty_like(u_d)

    def pure_copies(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            with torch.cuda.stream(s_in):
                scratch.copy_(u_h, non_blocking=True)
            with torch.cuda.stream(s_out):
                du_h.copy_(du_d, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3
    pure_copies(2)
    if dist is not None:
        dist.barrier()
    copy_ms = pure_copies(5)
    del scratch
    e2e_stats = eng.stats()
    e2e_sched = e2e_stats.get("host_schedule_rhs", "undecided")
    if dist is not None:
        tmax = torch.tensor([e2e_ms, e2e_plain_ms, copy_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_plain_ms, copy_ms = float(tmax[0].item()), float(tmax[1].item()), float(tmax[2].item())

    tplres = None
    st = eng.stats()
    if not args.no_template:
        eng.close()
        tplres = time_template(torch, G, p.prob, nx, args.steps, args.warmup, peak, lmesh, dist)
        if rank == 0:
            sys.stderr.write("[bench] template SpMV %.3f ms (tile kernel %.3f)  %.0f GB/s (%.2f of peak)  Tsit5 %.3f ms/step\n"
                             % (tplres["spmv_ms"], tplres["spmv_kernel_ms"], tplres["spmv_gbs"], tplres["spmv_frac"], tplres["tsit5_ms_per_step"]))
    extras = None
    if not args.no_extras and world == 1:
        extras = time_extras(torch, G, nx, max(10, args.steps // 4), args.warmup, peak)
        sys.stderr.write("[bench] extras: %s\n" % json.dumps(extras))
    strong = None
    if not args.no_strong:
        if eng is not None:
            eng.close()
        del u_d, du_d
        torch.cuda.empty_cache()
        strong = strong_scaling_leg(torch, G, dist, rank, world, local, max(10, args.steps // 2), args.warmup, peak, args.strong_nx,
                                    args.strong_partition)
        if rank == 0:
            sys.stderr.write("[bench] strong %s: RHS %.3f ms  SpMV %.3f ms  Tsit5 %.3f ms/step  parity %s\n"
                             % (strong["mesh"], strong["rhs_ms"], strong["spmv_ms"], strong["tsit5_ms_per_step"], strong["parity_max_rel"]))
    if rank != 0:
        dist.destroy_process_group()
        return
    cb = steady_cpu = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.ref_nx or nx)
        try:
            steady_cpu = cpu_steady_baseline(512)
        except Exception as e:
            steady_cpu = {"unavailable": type(e).__name__}
    traffic, traffic_src = profiled_traffic(args.variant) if (nx == 4096 and world == 1) else (None, None)
    launches_per_step = 1 + (1 if st["n_interface"] + (N - st["n_vertices"]) > 0 else 0) + (1 if st["n_live_boundary_edges"] else 0)
    line = {
        "metric": "fvm_eqs! Mtriangle-updates/s", "value": head["mtri_s"], "unit": "Mtriangle-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "README diffusion FVMProblem fvm_eqs! on triangulate_rectangle %dx%d per GPU (row strips of the "
                               "%dx%d lattice on [0,2]x[0,%d], Dirichlet u=0, D=1/9, u=50*U(0,1)); %d nodes, %d triangles per GPU; "
                               "one-layer node halo exchanged by NCCL send/recv before every RHS / SpMV when n_gpus > 1"
                               % (nx, nx, nx, nx * world, 2 * world, N, T),
                   "variant": args.variant, "l2": "inputs larger than L2 (%.2f GB streamed per step vs 126 MB L2)" % (head["alg_bytes"] / 1e9),
                   "tile_triangles": st["tile_triangles"], "n_tiles": st["n_tiles"]},
        "roofline": {"bound": "hbm", "kernel": "rhs_stream_kernel" if head["layout"] == "recompute" else "rhs_tile_kernel",
                     "achieved": head["kernel_gbs"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                     "frac_vs_minimal_bytes": head["frac_vs_minimal_bytes"],
                     "traffic": traffic, "traffic_source": ("from the committed ncu capture " + traffic_src + ", not measured in this run") if traffic else None,
                     "peak_source": peak_src,
                     "bytes_formula": {"general": "180*T + 25*N", "reduced": "108*T + 25*N", "recompute": "12*T + 41*N"}[head["layout"]],
                     "alg_bytes_per_launch": head["alg_bytes"], "kernel_ms": head["kernel_ms"]},
        "e2e": {"value": world * T / e2e_ms / 1e3, "unit": "Mtriangle-updates/s", "h2d_bytes_per_step": 8 * N,
                "d2h_bytes_per_step": 8 * N, "ms_per_step": e2e_ms,
                "schedule": ("banded pipeline, %d bands (copy-in / tiles / copy-out overlapped)" % e2e_stats["pipe_bands"])
                if e2e_sched == "pipeline" else "H2D, kernels, D2H in sequence",
                "schedule_choice": "one-off timing per handle (call 1 pipelined, call 2 plain, then the faster): " + e2e_sched,
                "unpipelined_ms_per_step": e2e_plain_ms,
                "host_link_copies_only_ms": copy_ms,
                "host_link_note": "the same two pinned buffers copied H2D and D2H on two streams with no kernel, all ranks at once "
                                  "(max over ranks): the floor this host's PCIe path sets for one call"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "setup_s": head["setup_s"],
    }
    if tplres:
        line["spmv"] = {"metric": "DiffusionEquation template y = A x + b, fp64 CSR SpMV", "gbs": tplres["spmv_gbs"],
                        "frac": tplres["spmv_frac"], "tile_kernel_ms": tplres["spmv_kernel_ms"], "ms_per_step": tplres["spmv_ms"],
                        "launches_per_spmv": 1, "traffic": profiled_traffic("spmv")[0] if (nx == 4096 and world == 1) else None,
                        "traffic_source": "from the committed ncu capture, not measured in this run", "format": "sliced ELL per tile, 16-bit tile-local columns, x staged in shared memory",
                        "alg_bytes_per_launch": tplres["spmv_alg_bytes"], "bytes_formula": "12*nnz + 4*(N+1) + 24*N",
                        "nnz": tplres["nnz"], "assemble_setup_s": tplres["assemble_setup_s"],
                        "e2e_ms_host_vectors": tplres["spmv_e2e_ms"]}
        line["tsit5"] = {"ms_per_step": tplres["tsit5_ms_per_step"], "steps": tplres["tsit5_steps"], "dt": tplres["tsit5_dt"],
                         "spmv_per_step": 6, "finite": tplres["finite"]}
        line["gpu_launches"] += args.steps + tplres["tsit5_launches"]
    if extras:
        line["other_configs"] = extras
    if cb:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_flat", "cpu_spmv_1t", "cpu_spmv_nt", "julia")
                                if k in cb}
        line["cpu_baseline"]["same_mesh_as_gpu_arm"] = cb["nx"] == nx
        if steady_cpu:
            line["cpu_baseline"]["cpu_steady_superlu"] = steady_cpu
    if halo_mode is not None:
        line["halo_exchange"] = {0: "none", 1: "grouped ncclSend/ncclRecv (pack, exchange, unpack)",
                                 2: "peer-mapped: one kernel stores the boundary values into the neighbours' slabs over NVLink "
                                    "(CUDA IPC), one kernel waits for their epoch flags and scatters"}[halo_mode]
    if parity_max_rel is not None:
        line["parity_max_rel"] = parity_max_rel
        line["parity_note"] = ("first / middle / last owned row of every rank's strip (the rows that read the NCCL-exchanged ghost layer) against a "
                               "single-domain evaluation of a 5-row patch of the global lattice on the same GPU; max over rows and ranks")
    if strong:
        line["strong_%d" % args.strong_nx] = strong
    if len(results) > 1:
        line["variants"] = {k: {kk: v[kk] for kk in ("ms_per_step", "mtri_s", "kernel_ms", "kernel_gbs", "frac", "frac_vs_minimal_bytes",
                                                     "step_frac_vs_minimal_bytes", "layout")}
                            for k, v in results.items()}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
