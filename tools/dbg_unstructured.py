import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import Pair, delaunay_mesh, rel_err
gtri = delaunay_mesh(200000, 77, jitter=0.35)
pair = Pair(gtri)
u = 0.2 + np.random.default_rng(3).random(gtri.num_points)
for flux, src in ((G.PowerDiffusion(0.3, 2.0), G.LogisticSource(1.3)), (G.ConstantDiffusion(0.7), None)):
    gp, op = pair.problem(G.Const(0.0), G.Neumann, flux, source=src)
    ref = O.fvm_eqs_vec(np.zeros_like(u), u, op, 0.0)
    # oracle with reversed triangle order: the spread that summation order alone produces
    otri2 = O.Triangulation(gtri.points, gtri.triangles[::-1].astype(np.int64), [np.asarray(s) for s in gtri.boundary_sections])
    for mode in (0, 1):
        p = G.get_cuda_parameters(gp, geometry_mode=mode)
        du = G.fvm_eqs(np.empty_like(u), u, p, 0.0)
        print(type(flux).__name__, mode, 'inf', rel_err(du, ref), 'l2', np.linalg.norm(du-ref)/np.linalg.norm(ref), 'max|ref|', np.abs(ref).max())
        p.engine.close()
