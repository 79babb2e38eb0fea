"""Experiment driver: the constant-D recompute kernel at 4096^2 over (resident CTAs, unroll, tile size);
checks every variant against the stored-geometry result and prints ms per variant."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = 4096
flux_f, gmode, layout = bench.VARIANTS["const_recompute"]
ref = None
u_h = None
for tt in (768, 1024, 1536):
    prob, _ = bench.lattice_problem(G, nx, nx, flux_f(G))
    p = G.get_cuda_parameters(prob, tile_triangles=tt, geometry_mode=gmode)
    eng = p.engine
    if u_h is None:
        u_h = 50.0 * torch.rand(eng.N, dtype=torch.float64)
    node_perm, _ = eng.permutation()
    u_d = u_h[torch.from_numpy(node_perm.astype(np.int64))].cuda()   # native order of this tiling
    du_d = torch.empty_like(u_d)
    for occ, unr in ((0, 1), (3, 1), (5, 1), (4, 2), (3, 2), (2, 2)):
        if occ == 5 and tt > 896:
            continue
        os.environ["FVM_REC_OCC"], os.environ["FVM_REC_UNR"] = str(occ), str(unr)
        ms, kms = bench.time_rhs(torch, eng, u_d, du_d, 40, 5)
        out = np.empty(eng.N)
        out[node_perm] = du_d.cpu().numpy()                            # back to caller order
        if ref is None:
            ref = out
        err = np.abs(out - ref).max() / np.abs(ref).max()
        print("const_recompute tile %4d occ %d unr %d  %.3f ms/step  kernel %.3f ms  maxdiff vs first %.1e" % (tt, occ, unr, ms, kms, err), flush=True)
    eng.close()
    del u_d, du_d, p, eng
