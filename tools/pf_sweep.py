"""Experiment driver: L2 prefetch distance (in tiles) of the recompute kernels at 4096^2."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = 4096
ref = {}
for name in ("const_recompute", "general_recompute"):
    flux_f, gmode, layout = bench.VARIANTS[name]
    for ahead in (0, 148, 296, 592, 888, 1184, 2368):
        os.environ["FVM_PF_AHEAD"] = str(ahead)
        prob, _ = bench.lattice_problem(G, nx, nx, flux_f(G))
        p = G.get_cuda_parameters(prob, geometry_mode=gmode)
        eng = p.engine
        torch.manual_seed(1)
        u_d = 50.0 * torch.rand(eng.N, dtype=torch.float64, device="cuda")
        du_d = torch.empty_like(u_d)
        ms, kms = bench.time_rhs(torch, eng, u_d, du_d, 60, 5)
        out = du_d.cpu().numpy()
        same = np.array_equal(out, ref.setdefault(name, out))
        print("%-18s ahead %4d  %.3f ms/step  kernel %.3f ms  identical=%s" % (name, ahead, ms, kms, same), flush=True)
        eng.close()
        del u_d, du_d, p, eng
