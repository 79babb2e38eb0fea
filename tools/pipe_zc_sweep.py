"""Experiment driver: host-buffer fvm_eqs! at 4096^2 with the bands moved by the copy engines (FVM_PIPE_ZC=0) or by
the zero-copy renumbering kernels (bit 0: copy-in, bit 1: copy-out), over band and CTA counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

os.environ["FVM_PIPE_AUTOTUNE"] = "0"
flux_f, gmode, layout = bench.VARIANTS["const_recompute"]
prob, _ = bench.lattice_problem(G, 4096, 4096, flux_f(G))
N = prob.mesh.triangulation.num_points
u_h = torch.empty(N, dtype=torch.float64).pin_memory()
du_h = torch.empty(N, dtype=torch.float64).pin_memory()
u_h.copy_(torch.from_numpy(50.0 * np.random.default_rng(1).random(N)))
un, dn = u_h.numpy(), du_h.numpy()
ref = None
for K in (10, 16, 24):
    os.environ["FVM_PIPE_BANDS"] = str(K)
    p = G.get_cuda_parameters(prob, geometry_mode=gmode)
    for zc, ctas in ((0, 64), (3, 32), (3, 64), (3, 148), (3, 296), (1, 64), (2, 64)):
        if K != 10 and (zc, ctas) not in ((0, 64), (3, 64), (3, 148)):
            continue
        os.environ["FVM_PIPE_ZC"], os.environ["FVM_PIPE_ZC_CTAS"] = str(zc), str(ctas)
        for _ in range(3):
            G.fvm_eqs(dn, un, p, 0.0)
        best = 1e9
        t00 = time.perf_counter()
        for _ in range(12):
            t0 = time.perf_counter()
            G.fvm_eqs(dn, un, p, 0.0)
            best = min(best, time.perf_counter() - t0)
        ms = (time.perf_counter() - t00) / 12 * 1e3
        if ref is None:
            ref = dn.copy()
        print("bands %2d  zc %d  ctas %3d   %.3f ms mean  %.3f ms best   identical=%s" % (K, zc, ctas, ms, best * 1e3, np.array_equal(dn, ref)), flush=True)
    p.engine.close()
    del p
