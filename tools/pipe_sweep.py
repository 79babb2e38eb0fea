"""Experiment driver: the host-buffer fvm_eqs! call (e2e) at 4096^2 over the pipeline's band layout.
Set PIPE_SWEEP_PAGEABLE=1 to also time pageable and fvm_host_register'ed NumPy arrays."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

nx = 4096
flux_f, gmode, layout = bench.VARIANTS["const_stored"]
ref = None
prob, _ = bench.lattice_problem(G, nx, nx, flux_f(G))
N = prob.mesh.triangulation.num_points
u_h = torch.empty(N, dtype=torch.float64).pin_memory()
du_h = torch.empty(N, dtype=torch.float64).pin_memory()
u_h.copy_(torch.from_numpy(50.0 * np.random.default_rng(1).random(N)))
un, dn = u_h.numpy(), du_h.numpy()
for K, taper in ((0, 0), (8, 0), (8, 1), (10, 1), (12, 1), (14, 1), (16, 1)):
    for k in ("FVM_NO_PIPELINE", "FVM_PIPE_BANDS", "FVM_PIPE_TAPER"):
        os.environ.pop(k, None)
    if K == 0:
        os.environ["FVM_NO_PIPELINE"] = "1"
    else:
        os.environ["FVM_PIPE_BANDS"], os.environ["FVM_PIPE_TAPER"] = str(K), str(taper)
    p = G.get_cuda_parameters(prob)
    for _ in range(2):
        G.fvm_eqs(dn, un, p, 0.0)
    t0 = time.perf_counter()
    for _ in range(10):
        G.fvm_eqs(dn, un, p, 0.0)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    if ref is None:
        ref = dn.copy()
    st = p.engine.stats()
    print("bands %2d taper %d  %.3f ms per host-buffer fvm_eqs!  identical=%s  early=%d calls=%d" % (K, taper, ms, np.array_equal(dn, ref), st["pipe_early_bands"], st["pipe_calls"]), flush=True)
    if os.environ.get("PIPE_SWEEP_PAGEABLE") and K in (0, 10):
        up, dp = un.copy(), np.empty(N)
        G.fvm_eqs(dp, up, p, 0.0)
        t0 = time.perf_counter()
        for _ in range(4):
            G.fvm_eqs(dp, up, p, 0.0)
        print("          pageable buffers: %.3f ms  identical=%s" % ((time.perf_counter() - t0) / 4 * 1e3, np.array_equal(dp, ref)), flush=True)
        with G.pinned(up, dp):   # the same arrays page-locked through fvm_host_register
            G.fvm_eqs(dp, up, p, 0.0)
            t0 = time.perf_counter()
            for _ in range(4):
                G.fvm_eqs(dp, up, p, 0.0)
            print("          registered buffers: %.3f ms  identical=%s" % ((time.perf_counter() - t0) / 4 * 1e3, np.array_equal(dp, ref)), flush=True)
    p.engine.close()
    del p
