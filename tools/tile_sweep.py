"""Tile-size sweep of the RHS tile kernels at 4096^2 (experiment driver; prints ms per variant/tile)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = 4096
for name, tiles in (("general_stored", (896, 1024, 1152, 1280)), ("const_stored", (384, 448, 512, 576, 640)), ("const_recompute", (896, 1024, 1152))):
    flux_f, gmode, layout = bench.VARIANTS[name]
    for tt in tiles:
        prob, _ = bench.lattice_problem(G, nx, nx, flux_f(G))
        p = G.get_cuda_parameters(prob, tile_triangles=tt, geometry_mode=gmode)
        eng = p.engine
        u_d = 50.0 * torch.rand(eng.N, dtype=torch.float64, device="cuda")
        du_d = torch.empty_like(u_d)
        ms, kms = bench.time_rhs(torch, eng, u_d, du_d, 60, 5)
        print("%-16s tile %4d  %.3f ms/step  kernel %.3f ms" % (name, tt, ms, kms), flush=True)
        eng.close()
        del u_d, du_d, p, eng
