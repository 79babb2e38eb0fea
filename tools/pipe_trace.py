"""Timeline of one host-buffer fvm_eqs! call at 4096^2 (FVM_PIPE_TRACE=1: per-band device timestamps on stderr):
    python tools/pipe_trace.py [bands] [trace 0/1]"""
import os, sys, time
K = sys.argv[1] if len(sys.argv) > 1 else "10"
os.environ["FVM_PIPE_BANDS"] = K
os.environ["FVM_PIPE_AUTOTUNE"] = "0"
if len(sys.argv) <= 2 or sys.argv[2] != "0":
    os.environ["FVM_PIPE_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

flux_f, gmode, layout = bench.VARIANTS["const_recompute"]
prob, _ = bench.lattice_problem(G, 4096, 4096, flux_f(G))
N = prob.mesh.triangulation.num_points
u_h = torch.empty(N, dtype=torch.float64).pin_memory()
du_h = torch.empty(N, dtype=torch.float64).pin_memory()
u_h.copy_(torch.from_numpy(50.0 * np.random.default_rng(1).random(N)))
un, dn = u_h.numpy(), du_h.numpy()
p = G.get_cuda_parameters(prob, geometry_mode=gmode)
for i in range(6):
    sys.stderr.write("---- call %d\n" % i)
    t0 = time.perf_counter()
    G.fvm_eqs(dn, un, p, 0.0)
    sys.stderr.write("     wall %.3f ms\n" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
for _ in range(10):
    G.fvm_eqs(dn, un, p, 0.0)
print("bands %s: %.3f ms per host-buffer fvm_eqs! (10 calls)" % (K, (time.perf_counter() - t0) / 10 * 1e3), p.engine.stats(), flush=True)
