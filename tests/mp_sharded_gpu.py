"""Multi-rank GPU parity (run under torchrun, one rank per GPU): the sharded fvm_eqs!, template
SpMV and device Tsit5 with NCCL halo exchange must reproduce the single-domain ORACLE on the owned
nodes of every rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import fvm_b200 as G
    from oracle import fvm_oracle as O
    from tests.common import RTOL_RHS, RTOL_TSIT5, rel_err
    from tests.sharding_worker import build_case
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("FVM_HALO_OVERLAP", "1")  # exercise the overlapped schedule (opt-in) in the parity run
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    errs = {}

    def log(msg):
        sys.stderr.write("[rank %d] %s\n" % (rank, msg))
        sys.stderr.flush()

    for kind in ("lattice", "delaunay"):
        log("case " + kind)
        pair, gp, op, owner = build_case(kind)
        if world != 2:
            owner = G.partition_rcb(pair.gtri.points, world)
        local = G.extract_local(pair.gtri, owner, rank, world)
        lp = G.shard_problem(gp, local)
        p = G.get_sharded_cuda_parameters(lp, local, dist, tile_triangles=128, device=local_rank)
        N = pair.gtri.num_points
        u = 0.2 + np.random.default_rng(5).random(N)
        ref = O.fvm_eqs_vec(np.zeros(N), u, op, 0.3)
        ul = u[local.global_nodes].copy()
        ul[local.n_owned:] = 1e300  # ghosts must come from the NCCL exchange
        log("halo installed")
        du = G.fvm_eqs(np.zeros_like(ul), ul, p, 0.3)
        log("rhs done")
        own = local.global_nodes[:local.n_owned]
        errs["rhs_" + kind] = rel_err(du[:local.n_owned], ref[own]) if np.abs(ref).max() > 0 else 0.0
        assert errs["rhs_" + kind] <= RTOL_RHS, errs
        # the ghost refresh runs through peer-mapped NVLink stores when CUDA IPC is available, else through NCCL
        # send/recv; both give the same bits
        mode, timed_out = p.engine.halo_mode()
        log("halo exchange mode %d (1 = NCCL send/recv, 2 = peer-mapped stores)" % mode)
        assert mode in (1, 2) and timed_out == 0
        os.environ["FVM_HALO_PEER"] = "0"
        p3 = G.get_sharded_cuda_parameters(lp, local, dist, tile_triangles=128, device=local_rank)
        del os.environ["FVM_HALO_PEER"]
        assert p3.engine.halo_mode()[0] == 1
        du3 = G.fvm_eqs(np.zeros_like(ul), ul, p3, 0.3)
        assert np.array_equal(du3, du), np.abs(du3 - du).max()
        for rep in range(20):  # many exchanges back to back: the two slab parities and the epoch flags
            du4 = G.fvm_eqs(np.zeros_like(ul), ul * (1.0 + 0.01 * rep), p, 0.3)
            du5 = G.fvm_eqs(np.zeros_like(ul), ul * (1.0 + 0.01 * rep), p3, 0.3)
            assert np.array_equal(du4, du5), rep
        p3.engine.close()
        errs["halo_mode_" + kind] = float(mode)
        # the banded host-buffer pipeline (fvm_pipe.cu) with the halo exchange at the head of its last stage:
        # forced on for this small mesh, bit-identical to the plain schedule on every rank
        for bands in (3, 6):
            os.environ.update(FVM_PIPE_MIN_NODES="0", FVM_PIPE_FORCE="1", FVM_PIPE_BANDS=str(bands))
            p2 = G.get_sharded_cuda_parameters(lp, local, dist, tile_triangles=128, device=local_rank)
            du2 = G.fvm_eqs(np.full_like(ul, np.nan), ul, p2, 0.3)
            assert p2.engine.stats()["pipe_calls"] == 1, p2.engine.stats()
            assert np.array_equal(du2, du), (kind, bands, np.abs(du2 - du).max())
            p2.engine.close()
            for k in ("FVM_PIPE_MIN_NODES", "FVM_PIPE_FORCE", "FVM_PIPE_BANDS"):
                del os.environ[k]
        log("pipelined rhs identical")
        # device Tsit5 on the sharded RHS vs the single-domain oracle
        if kind == "delaunay":
            dt, t1 = 2e-5, 2e-4
            ul = gp.initial_condition[local.global_nodes] + 0.3 + 0.1 * np.sin(7 * pair.gtri.points[local.global_nodes, 0])
            ug = gp.initial_condition + 0.3 + 0.1 * np.sin(7 * pair.gtri.points[:, 0])
            from fvm_b200 import _lib as L
            ul = np.ascontiguousarray(ul)
            L.check(p.engine.h, L.lib().fvm_tsit5(p.engine.h, 0, ul.ctypes.data, 0.0, t1, dt, 0, None, None, 0))
            uref = O.tsit5_fixed(lambda d, x, t: O.fvm_eqs_vec(d, x, op, t), ug, 0.0, t1, dt,
                                 callback=lambda x, t: (O.update_dirichlet_nodes(x, t, op), True)[1])
            errs["tsit5_" + kind] = rel_err(ul[:local.n_owned], uref[own])
            assert errs["tsit5_" + kind] <= RTOL_TSIT5, errs
        p.engine.close()
    # template operator: sharded SpMV y = A x + b and Tsit5 on the lattice
    tri = G.triangulate_rectangle(0, 2, 0, 2, 40, 32, single_boundary=True)
    owner = G.partition_strips(tri.points, world)
    local = G.extract_local(tri, owner, rank, world)
    lmesh = G.FVMGeometry(local.triangulation)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    tpl = G.DiffusionEquation(lmesh, G.BoundaryConditions(lmesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9,
                              initial_condition=ic[local.global_nodes], final_time=0.02, ghost=local.is_ghost)
    log("template assembled")
    G.install_halo(tpl.engine, local, dist)
    otri = O.triangulate_rectangle(0, 2, 0, 2, 40, 32, single_boundary=True)
    omesh = O.FVMGeometry(otri)
    ref = O.DiffusionEquation(omesh, O.BoundaryConditions(omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet),
                              diffusion_function=lambda x, y, p: 1 / 9, initial_condition=ic, final_time=0.02)
    x = np.random.default_rng(9).random(tri.num_points)
    xl = x[local.global_nodes].copy()
    xl[local.n_owned:] = 1e300
    y = tpl.mul(np.empty_like(xl), xl)
    own = local.global_nodes[:local.n_owned]
    errs["spmv"] = rel_err(y[:local.n_owned], (ref.A @ x + ref.b)[own])
    assert errs["spmv"] <= RTOL_RHS, errs
    os.environ.update(FVM_PIPE_MIN_NODES="0", FVM_PIPE_FORCE="1", FVM_PIPE_BANDS="4")
    tpl2 = G.DiffusionEquation(lmesh, G.BoundaryConditions(lmesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9,
                               initial_condition=ic[local.global_nodes], final_time=0.02, ghost=local.is_ghost, tile_triangles=128)
    G.install_halo(tpl2.engine, local, dist)
    y2 = tpl2.mul(np.full_like(xl, np.nan), xl)
    assert tpl2.engine.stats()["pipe_calls"] == 1 and rel_err(y2[:local.n_owned], y[:local.n_owned]) <= 1e-14, np.abs(y2 - y).max()
    tpl2.engine.close()
    for k in ("FVM_PIPE_MIN_NODES", "FVM_PIPE_FORCE", "FVM_PIPE_BANDS"):
        del os.environ[k]
    sol = G.solve(tpl, G.Tsit5(0.001))
    uref = O.tsit5_fixed(lambda d, v, t: d.__setitem__(Ellipsis, ref.A @ v + ref.b), ref.u0, 0.0, 0.02, 0.001)
    errs["tsit5_operator"] = rel_err(sol.u[:local.n_owned], uref[own])
    assert errs["tsit5_operator"] <= RTOL_TSIT5, errs
    tpl.engine.close()
    # sharded steady solve: Poisson on the lattice, PCG and BiCGStab with NCCL all-reduced dot products
    log("sharded krylov")
    tri = G.triangulate_rectangle(0, 1, 0, 1, 60, 48, single_boundary=True)
    owner = G.partition_rcb(tri.points, world)
    local = G.extract_local(tri, owner, rank, world)
    lmesh = G.FVMGeometry(local.triangulation)
    src = lambda x, y, p: -np.sin(np.pi * x) * np.sin(np.pi * y)
    otri = O.triangulate_rectangle(0, 1, 0, 1, 60, 48, single_boundary=True)
    omesh = O.FVMGeometry(otri)
    ref = O.PoissonsEquation(omesh, O.BoundaryConditions(omesh, lambda x, y, t, u, p: 0.0 * x, O.Dirichlet), source_function=src)
    uref = O.solve_steady(ref)
    own = local.global_nodes[:local.n_owned]
    for method in ("pcg", "bicgstab"):
        tpl = G.PoissonsEquation(lmesh, G.BoundaryConditions(lmesh, G.Const(0.0), G.Dirichlet), source_function=src,
                                 ghost=local.is_ghost, tile_triangles=256)
        G.install_halo(tpl.engine, local, dist)
        sol = G.solve(tpl, G.KrylovJacobi(method, rtol=1e-13))
        errs["krylov_" + method] = rel_err(sol.u[:local.n_owned], uref[own])
        assert errs["krylov_" + method] <= 1e-9 and sol.relres <= 1e-11, (errs, sol.relres, sol.iters)
        tpl.engine.close()
    # a problem only ONE rank can see is invalid (a Dudt node inside rank 0's subdomain, steady template) is rejected by
    # every rank with the same error, before any collective of the solver is entered
    log("rank-consistent validation")
    far = int(np.argmin(np.abs(tri.points - np.array([0.5, 0.5])).sum(axis=1)))
    dud = {int(np.nonzero(local.global_nodes == far)[0][0]): 0} if far in set(local.global_nodes[:local.n_owned].tolist()) else {}
    ics = G.InternalConditions((G.Const(1.0),), dudt_nodes=dud)
    try:
        bad = G.PoissonsEquation(lmesh, G.BoundaryConditions(lmesh, G.Const(0.0), G.Dirichlet), ics, source_function=src,
                                 ghost=local.is_ghost, tile_triangles=256)
        G.install_halo(bad.engine, local, dist)
        raise AssertionError("rank %d accepted a problem another rank rejects" % rank)
    except ValueError as e:
        assert "does not support Dudt nodes" in str(e), str(e)
    # adaptive Tsit5: ghost entries are excluded from the error norm, so the accepted / rejected step counts of the
    # sharded run equal those of a single-domain run
    log("adaptive step counts")
    tri = G.triangulate_rectangle(0, 2, 0, 2, 40, 32, single_boundary=True)
    owner = G.partition_strips(tri.points, world)
    local = G.extract_local(tri, owner, rank, world)
    lmesh = G.FVMGeometry(local.triangulation)
    ic = np.where(tri.points[:, 1] <= 1.0, 50.0, 0.0)
    tpl = G.DiffusionEquation(lmesh, G.BoundaryConditions(lmesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9,
                              initial_condition=ic[local.global_nodes], final_time=0.05, ghost=local.is_ghost)
    G.install_halo(tpl.engine, local, dist)
    a_sh, a_one = G.Tsit5(), G.Tsit5()
    sol = G.solve(tpl, a_sh)
    gmesh = G.FVMGeometry(tri)
    one = G.DiffusionEquation(gmesh, G.BoundaryConditions(gmesh, G.Const(0.0), G.Dirichlet), diffusion_function=1 / 9,
                              initial_condition=ic, final_time=0.05)
    sol1 = G.solve(one, a_one)
    assert (a_sh.naccept, a_sh.nreject) == (a_one.naccept, a_one.nreject), ((a_sh.naccept, a_sh.nreject), (a_one.naccept, a_one.nreject))
    errs["adaptive_vs_single"] = rel_err(sol.u[:local.n_owned], sol1.u[local.global_nodes[:local.n_owned]])
    assert errs["adaptive_vs_single"] <= 1e-10, errs
    tpl.engine.close()
    one.engine.close()
    dist.barrier()
    print("rank %d/%d sharded parity OK: %s" % (rank, world, {k: float("%.2e" % v) for k, v in errs.items()}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failing rank must not hang its peers inside an NCCL call
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    os._exit(0)
