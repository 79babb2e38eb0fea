"""Experiment driver: host-buffer fvm_eqs! at 4096^2 over the pipeline's band count and taper (copy engines)."""
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
combos = [(10, 1), (8, 1), (8, 2), (10, 2), (12, 1), (12, 2), (14, 2), (16, 2), (10, 1)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for K, taper in combos:
    os.environ["FVM_PIPE_BANDS"], os.environ["FVM_PIPE_TAPER"] = str(K), str(taper)
    p = G.get_cuda_parameters(prob, geometry_mode=gmode)
    for _ in range(3):
        G.fvm_eqs(dn, un, p, 0.0)
    best = 1e9
    t00 = time.perf_counter()
    for _ in range(15):
        t0 = time.perf_counter()
        G.fvm_eqs(dn, un, p, 0.0)
        best = min(best, time.perf_counter() - t0)
    ms = (time.perf_counter() - t00) / 15 * 1e3
    if ref is None:
        ref = dn.copy()
    st = p.engine.stats()
    print("bands %2d  taper %d   %.3f ms mean  %.3f ms best   identical=%s  early=%d" % (K, taper, ms, best * 1e3, np.array_equal(dn, ref), st["pipe_early_bands"]), flush=True)
    p.engine.close()
    del p
