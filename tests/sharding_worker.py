"""World-size-2 worker for the sharding tests (gloo on CPU): partitions a mesh, exchanges the
ghost layer according to the halo plan and evaluates the rank-local ORACLE right-hand side."""
import os

import numpy as np


def oracle_local_problem(O, oprob, local):
    """Rank-local oracle FVMProblem: same closures, dict conditions mapped to local ids."""
    lt = local.triangulation
    otri = O.Triangulation(lt.points, lt.triangles.astype(np.int64), boundary_sections=[])
    uv, sec = lt.boundary_edges()
    for k, ((u, v), s) in enumerate(zip(uv.tolist(), sec.tolist())):
        otri.boundary_edge_map[(u, v)] = (s, k)
        otri.adjacent[(v, u)] = -(s + 1)
    omesh = O.FVMGeometry(otri)
    g2l = {int(g): i for i, g in enumerate(local.global_nodes)}
    gc = oprob.conditions
    c = O.Conditions.__new__(O.Conditions)
    c.functions = gc.functions
    c.dirichlet_nodes = {g2l[n]: f for n, f in gc.dirichlet_nodes.items() if n in g2l}
    c.dudt_nodes = {g2l[n]: f for n, f in gc.dudt_nodes.items() if n in g2l}
    c.neumann_edges = {(g2l[a], g2l[b]): f for (a, b), f in gc.neumann_edges.items()
                       if a in g2l and b in g2l and (g2l[a], g2l[b]) in otri.boundary_edge_map}
    c.constrained_edges = {(g2l[a], g2l[b]): f for (a, b), f in gc.constrained_edges.items()
                           if a in g2l and b in g2l and (g2l[a], g2l[b]) in otri.boundary_edge_map}
    p = O.FVMProblem.__new__(O.FVMProblem)
    p.__dict__.update(oprob.__dict__)
    p.mesh, p.conditions = omesh, c
    p.initial_condition = oprob.initial_condition[local.global_nodes]
    return p


def build_case(kind):
    import fvm_b200 as G
    from tests.common import Pair, delaunay_mesh
    from tests.test_rhs_gpu import _split_loop
    if kind == "lattice":
        pair = Pair(G.triangulate_rectangle(0, 2, 0, 2, 21, 16, single_boundary=True))
        gp, op = pair.problem(G.Const(0.0), G.Dirichlet, G.ConstantDiffusion(1 / 9))
        owner = G.partition_strips(pair.gtri.points, 2)
    else:
        pair = Pair(_split_loop(delaunay_mesh(500, 17, extra_points=3)))
        specs = (G.Const(0.25), G.AffineU(0.1, -0.5), G.LinearXY(0.3, 0.2, -0.1), G.Const(0.0))
        types = (G.Dirichlet, G.Dudt, G.Neumann, G.Constrained)
        internal = ((G.Const(0.7),), {200: 0}, {300: 0})
        gp, op = pair.problem(specs, types, G.PowerDiffusion(0.3, 2.0), source=G.LogisticSource(1.3), internal=internal)
        # "graph": the METIS-style partitioner of the library (irregular interface); else coordinate bisection
        owner = G.partition_graph(pair.gtri, 2) if kind == "graph" else G.partition_rcb(pair.gtri.points, 2)
    return pair, gp, op, owner


def exchange_host(dist, torch, u_local, local):
    """the halo exchange schedule of fvm_halo_exchange, on host tensors over gloo"""
    reqs, bufs = [], []
    for q, snd, rcv in zip(local.neighbours, local.send_nodes, local.recv_nodes):
        sb = torch.from_numpy(np.ascontiguousarray(u_local[snd]))
        rb = torch.empty(len(rcv), dtype=torch.float64)
        reqs.append(dist.isend(sb, q))
        reqs.append(dist.irecv(rb, q))
        bufs.append((rcv, rb, sb))
    for r in reqs:
        r.wait()
    for rcv, rb, _ in bufs:
        u_local[rcv] = rb.numpy()


def worker(rank, world, port, kind, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import fvm_b200 as G
    from oracle import fvm_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pair, gp, op, owner = build_case(kind)
        local = G.extract_local(pair.gtri, owner, rank, world)
        N = pair.gtri.num_points
        u = 0.2 + np.random.default_rng(5).random(N)
        u_local = u[local.global_nodes].copy()
        u_local[local.n_owned:] = np.nan  # ghosts must come from the exchange
        exchange_host(dist, torch, u_local, local)
        assert np.array_equal(u_local, u[local.global_nodes])
        lp = oracle_local_problem(O, op, local)
        du_local = O.fvm_eqs_vec(np.zeros_like(u_local), u_local, lp, 0.3)
        # host-side mirror builds the same local problem (checked on the GPU by the multi-rank test)
        gl = G.shard_problem(gp, local)
        assert gl.conditions.node_kind.shape == (len(local.global_nodes),)
        np.save(os.path.join(out_dir, "du_%s_%d.npy" % (kind, rank)), du_local[:local.n_owned])
        np.save(os.path.join(out_dir, "own_%s_%d.npy" % (kind, rank)), local.global_nodes[:local.n_owned])
    finally:
        dist.destroy_process_group()
