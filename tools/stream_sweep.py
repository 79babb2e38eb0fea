"""Experiment driver: tile size x consumer threads of the streaming recompute kernel (fvm_rhs_stream.cu) at 4096^2.
Checks every configuration against the stored-geometry kernel's du (1e-12) before timing it."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = int(os.environ.get("SWEEP_NX", "4096"))
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["const_recompute"]
tiles = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "256,384,512,768,1024".split(","))]
threads = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "256,384,512".split(","))]  # 2564 = 256 threads, 4 CTAs/SM register budget
def system_problem():
    """BASELINE config 4: 2-species Keller-Segel FVMSystem, all-Neumann zero flux."""
    tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, nx, nx, single_boundary=True)
    mesh = G.FVMGeometry(tri)
    N = tri.num_points
    ks, kss = G.KellerSegelFlux(4.0, 1.0), G.KellerSegelSource(0.1)
    BC = G.BoundaryConditions(mesh, G.Const(0.0), G.Neumann)
    pu = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=np.zeros(N), final_time=1.0)
    pv = G.FVMProblem(mesh, BC, flux_function=ks, source_function=kss, initial_condition=np.zeros(N), final_time=1.0)
    return G.FVMSystem(pu, pv)


for name in names:
    torch.manual_seed(1)
    if name == "system":
        prob, gmode = system_problem(), 1
        u_d = 0.01 * torch.rand(2 * prob.mesh.triangulation.num_points, dtype=torch.float64, device="cuda")
    else:
        flux_f, gmode, layout = bench.VARIANTS[name]
        prob, _ = bench.lattice_problem(G, nx, nx, flux_f(G))
        u_d = 50.0 * torch.rand(prob.mesh.triangulation.num_points, dtype=torch.float64, device="cuda")
    p = G.get_cuda_parameters(prob, geometry_mode=0)
    du_ref = torch.empty_like(u_d)
    ms, kms = bench.time_rhs(torch, p.engine, u_d, du_ref, 30, 3)
    print("%-18s stored-geometry reference: %.3f ms/step" % (name, ms), flush=True)
    # native orders differ between tile sizes: compare in caller order
    p.engine.rhs_device(du_ref.data_ptr(), u_d.data_ptr(), 0.0, native=False)
    p.engine.close()
    for tt in tiles:
        for th in threads:
            os.environ["FVM_STREAM_THREADS"] = str(th if th not in (2564, 2562) else 256)  # 2564 / 2562: 256 threads, register budget for 4 / 2 CTAs per SM
            os.environ["FVM_STREAM_OCC"] = "4" if th == 2564 else ("2" if th == 2562 else "3")
            try:
                p = G.get_cuda_parameters(prob, tile_triangles=tt, geometry_mode=gmode)
            except Exception as e:
                print("%-18s TT %4d thr %3d  setup failed: %s" % (name, tt, th, e), flush=True)
                continue
            eng = p.engine
            du_d = torch.empty_like(u_d)
            try:
                eng.rhs_device(du_d.data_ptr(), u_d.data_ptr(), 0.0, native=False)
                err = float((du_d - du_ref).abs().max() / du_ref.abs().max())
                ms, kms = bench.time_rhs(torch, eng, u_d, du_d, 60, 5)
                print("%-18s TT %4d thr %3d  %.3f ms/step  kernel %.3f ms  rel err vs stored %.2e" % (name, tt, th, ms, kms, err), flush=True)
            except Exception as e:
                print("%-18s TT %4d thr %3d  failed: %s" % (name, tt, th, e), flush=True)
            eng.close()
            del du_d, p, eng
