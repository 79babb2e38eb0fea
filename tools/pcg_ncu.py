"""ncu launch-list target: a few Jacobi-PCG iterations on the MeanExitTimeProblem (2048^2 by default), graphs off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FVM_NO_GRAPH"] = "1"
import fvm_b200 as G
n3 = int(os.environ.get("PROF_N3", "2048"))
tri = G.triangulate_rectangle(0.0, 2.0, 0.0, 2.0, n3, n3, single_boundary=True)
mesh = G.FVMGeometry(tri)
met = G.MeanExitTimeProblem(mesh, G.BoundaryConditions(mesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9)
sol = G.solve(met, G.KrylovJacobi("pcg", rtol=1e-10, maxiter=int(os.environ.get("PROF_ITERS", "48"))))
print(sol.iters, sol.relres)
