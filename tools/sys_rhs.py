"""Experiment driver: FVMSystem (Keller-Segel, 2 species) RHS throughput at nx^2 for the current env knobs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True)
mesh = G.FVMGeometry(tri)
N = tri.num_points
ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann)
pu = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=np.zeros(N), final_time=1.0)
pv = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=np.zeros(N), final_time=1.0)
p = G.get_cuda_parameters(G.FVMSystem(pu, pv))
eng = p.engine
u_d = 0.01 * torch.rand(2 * N, dtype=torch.float64, device="cuda")
du_d = torch.empty_like(u_d)
ms, kms = bench.time_rhs(torch, eng, u_d, du_d, 50, 5)
B = 180 * eng.T + 42 * N
print("tile=%s minb=%s  %.3f ms/step  tile-kernel %.3f ms  %.1f Gtri/s  %.0f GB/s" % (
    os.environ.get("FVM_TILE_TRIANGLES", "default"), os.environ.get("FVM_SYS_MINB", "3"), ms, kms, eng.T / ms / 1e6, B / kms / 1e6))
