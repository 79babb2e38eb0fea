"""Experiment driver: time the device solvers under the switches of fvm_solvers.cu.
    python tools/solver_profile.py [pcg|tsit5|both]
(FVM_NO_FUSE=1: the unfused PCG iteration; Tsit5 has no fused form, both lines time the same code.)
PCG: MeanExitTimeProblem 2048^2 (BASELINE config 3), ms per iteration.  Tsit5: DiffusionEquation 4096^2, ms per step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G

what = sys.argv[1] if len(sys.argv) > 1 else "both"
variants = [("default", {}), ("nomerge", {"FVM_PCG_NO_MERGE": "1"}), ("nofuse", {"FVM_NO_FUSE": "1"}), ("default", {}), ("nomerge", {"FVM_PCG_NO_MERGE": "1"})]


def with_env(env, fn):
    for k, v in env.items():
        os.environ[k] = v
    try:
        return fn()
    finally:
        for k in env:
            del os.environ[k]


if what in ("pcg", "both"):
    n3 = int(os.environ.get("PROF_N3", "2048"))
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, n3, n3, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    met = G.MeanExitTimeProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9)
    for name, env in variants:

        def run():
            G.solve(met, G.KrylovJacobi("pcg", rtol=1e-10, maxiter=600))
            t0 = time.perf_counter()
            sol = G.solve(met, G.KrylovJacobi("pcg", rtol=1e-10, maxiter=40000))
            return sol, time.perf_counter() - t0
        sol, dt = with_env(env, run)
        print("pcg   %-8s %5d iters  %.4f ms/iter  relres %.2e" % (name, sol.iters, 1e3 * dt / sol.iters, sol.relres), flush=True)
    met.engine.close()

for nx in ([int(x) for x in os.environ.get("PROF_NX", "50,512,2048,4096").split(",")] if what in ("tsit5", "both") else []):
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    tpl = G.DiffusionEquation(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9, initial_condition=ic, final_time=1.0)
    h = 2.0 / (nx - 1)
    dt = 0.2 * 3.3 * h * h / (8.0 / 9.0)
    eng = tpl.engine
    ref = None
    nst = 40 if nx >= 2048 else 400
    for name, env in variants[:1]:

        def run():
            u = torch.from_numpy(tpl.u0).cuda()
            eng.tsit5_device(u.data_ptr(), 0.0, 2 * dt, dt, True)
            torch.cuda.synchronize()
            u = torch.from_numpy(tpl.u0).cuda()
            t0 = time.perf_counter()
            eng.tsit5_device(u.data_ptr(), 0.0, nst * dt, dt, True)
            torch.cuda.synchronize()
            return u.cpu().numpy(), (time.perf_counter() - t0) / nst
        u, per = with_env(env, run)
        if ref is None:
            ref = u
        print("tsit5 %5d %-8s %.4f ms/step  max diff vs default %.2e" % (nx, name, 1e3 * per, np.abs(u - ref).max()), flush=True)
    eng.close()
