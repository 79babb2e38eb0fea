"""Wall time of the phases of fvm_finalize / template assembly at nx^2 (FVM_TIMING=1 prints the C++ phases)."""
import os, sys, time
os.environ["FVM_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fvm_b200 as G
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = time.perf_counter(); tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True); t1 = time.perf_counter()
mesh = G.FVMGeometry(tri); t2 = time.perf_counter()
BCs = G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet)
ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
prob = G.FVMProblem(mesh, BCs, diffusion_function=G.ConstantDiffusion(1 / 9), initial_condition=ic, final_time=0.5); t3 = time.perf_counter()
p = G.get_cuda_parameters(prob); t4 = time.perf_counter()
print("python: triangulate %.3f  FVMGeometry %.3f  conditions+problem %.3f  get_cuda_parameters %.3f s" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3), flush=True)
p.engine.close()
t0 = time.perf_counter()
tpl = G.DiffusionEquation(mesh, BCs, diffusion_function=1 / 9, initial_condition=ic, final_time=1.0)
print("python: DiffusionEquation (finalize + assemble) %.3f s" % (time.perf_counter() - t0), flush=True)
