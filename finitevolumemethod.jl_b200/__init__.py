"""finitevolumemethod.jl_b200 — host-side mirror of FiniteVolumeMethod.jl's public API for the
B200-native `fvm_eqs!` / linear-template hot path (libfvmcuda.so).  Import it as `fvm_b200`
(see fvm_b200.py at the repository root; the directory name is not a valid Python identifier)."""
from ._lib import FVMCudaError, UnsupportedClosureError, LIB_PATH, exported_symbols  # noqa: F401
from .conditions import (BoundaryConditions, Conditions, Constrained, Dirichlet, Dudt,  # noqa: F401
                         InternalConditions, Neumann)
from .functors import *  # noqa: F401,F403
from .mesh import Triangulation, triangulate_rectangle  # noqa: F401
from .problem import (CudaParameters, Engine, InvalidFluxError, compute_flux, pl_interpolate, FVMGeometry, FVMProblem, FVMSystem,  # noqa: F401
                      SteadyFVMProblem, fvm_eqs, get_cuda_parameters, jacobian, jacobian_sparsity, pinned,
                      update_dirichlet_nodes)
from .templates import (DiffusionEquation, KrylovJacobi, LaplacesEquation,  # noqa: F401
                        LinearReactionDiffusionEquation, MeanExitTimeProblem, PoissonsEquation, Solution, Tsit5)
from .solve import NewtonRaphson, solve  # noqa: F401
from .sharding import (LocalMesh, edge_cut, extract_local, partition_graph, get_sharded_cuda_parameters, install_halo,  # noqa: F401
                       lattice_rows, lattice_strip_local, patch_mesh, partition_rcb, partition_strips, shard_problem)
from .wire import WireError, WireReader, WireWriter, load_mesh, load_solution, save_mesh, save_solution  # noqa: F401
