"""Runs one RHS variant of bench.py a few times at 4096^2 (driver for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:rhs_tile_kernel --launch-skip 3 -c 1 \\
        -o gpurun_out/prof python tools/run_variant.py const_recompute 6"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fvm_b200 as G
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "const_recompute"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
flux_f, gmode, layout = bench.VARIANTS[name]
prob, _ = bench.lattice_problem(G, 4096, 4096, flux_f(G))
p = G.get_cuda_parameters(prob, geometry_mode=gmode)
eng = p.engine
u_d = 50.0 * torch.rand(eng.N, dtype=torch.float64, device="cuda")
du_d = torch.empty_like(u_d)
for _ in range(steps):
    eng.rhs_device(du_d.data_ptr(), u_d.data_ptr(), 0.0, native=True)
eng.synchronize()
print(name, "done", eng.stats())
