"""N > 1 path on CPU (gloo, world_size 2): partitioners, rank-local meshes, halo plan and the
exchange schedule; the sharded ORACLE right-hand side must equal the single-domain one on every
owned node (only the summation order differs)."""
import os
import socket

import numpy as np
import pytest

import fvm_b200 as G
from oracle import fvm_oracle as O
from tests.common import rel_err
from tests.sharding_worker import build_case, worker


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind", ["lattice", "delaunay", "graph"])
def test_sharded_oracle_rhs_two_ranks(kind, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(2, _free_port(), kind, str(tmp_path)), nprocs=2, join=True)
    pair, gp, op, owner = build_case(kind)
    N = pair.gtri.num_points
    u = 0.2 + np.random.default_rng(5).random(N)
    ref = O.fvm_eqs_vec(np.zeros(N), u, op, 0.3)
    got = np.full(N, np.nan)
    for r in range(2):
        got[np.load(tmp_path / ("own_%s_%d.npy" % (kind, r)))] = np.load(tmp_path / ("du_%s_%d.npy" % (kind, r)))
    assert not np.isnan(got).any()  # every node is owned by exactly one rank
    assert rel_err(got, ref) <= 1e-13


def test_partitioners_and_local_meshes():
    tri = G.triangulate_rectangle(0, 2, 0, 2, 33, 32, single_boundary=True)
    for nparts in (2, 4, 8):
        for owner in (G.partition_strips(tri.points, nparts), G.partition_rcb(tri.points, nparts), G.partition_graph(tri, nparts)):
            cnt = np.bincount(owner, minlength=nparts)
            assert cnt.min() >= len(owner) // nparts - 1 and cnt.max() <= len(owner) // nparts + 1
            locs = [G.extract_local(tri, owner, r, nparts) for r in range(nparts)]
            # every triangle touching an owned node is local; ghosts are exactly the non-owned vertices
            for r, lm in enumerate(locs):
                gt = lm.global_nodes[lm.triangulation.triangles]
                assert ((owner[gt] == r).any(axis=1)).all()
                assert (owner[lm.global_nodes[:lm.n_owned]] == r).all() and (owner[lm.global_nodes[lm.n_owned:]] != r).all()
                # send/recv lists pair up across ranks in the same (global id) order
                for q, snd in zip(lm.neighbours, lm.send_nodes):
                    other = locs[q]
                    rcv = other.recv_nodes[other.neighbours.index(r)]
                    assert np.array_equal(lm.global_nodes[snd], other.global_nodes[rcv])
            assert sum(lm.n_owned for lm in locs) == tri.num_points
    # analytic strips used by bench.py equal the generic extraction
    owner = G.partition_strips(tri.points, 4)
    for r in range(4):
        a, b = G.lattice_strip_local(0, 2, 0, 2, 33, 8, r, 4), G.extract_local(tri, owner, r, 4)
        assert np.array_equal(a.global_nodes, b.global_nodes) and np.array_equal(a.triangulation.points, b.triangulation.points)
        assert sorted(map(tuple, a.triangulation.triangles.tolist())) == sorted(map(tuple, b.triangulation.triangles.tolist()))
        assert sorted(map(tuple, a.triangulation.boundary_edges()[0].tolist())) == sorted(map(tuple, b.triangulation.boundary_edges()[0].tolist()))
        assert a.neighbours == b.neighbours
        assert all(np.array_equal(x, y) for x, y in zip(a.send_nodes + a.recv_nodes, b.send_nodes + b.recv_nodes))


def test_graph_partitioner_balance_and_cut():
    """fvm_partition_graph (METIS-style recursive bisection: graph growing + FM refinement, host-only): equal part
    sizes for any k, every node assigned, an edge cut of the order of the geometric partitioners', and a partition
    that follows the connectivity where coordinates mislead (two blocks joined by a thin bridge)."""
    from tests.common import delaunay_mesh
    for tri in (G.triangulate_rectangle(0, 2, 0, 1, 80, 41, single_boundary=True), delaunay_mesh(6000, 11, jitter=0.3, extra_points=3)):
        for k in (2, 3, 5, 8):
            owner = G.partition_graph(tri, k)
            cnt = np.bincount(owner, minlength=k)
            assert owner.min() == 0 and owner.max() == k - 1 and cnt.max() - cnt.min() <= 1
            assert G.edge_cut(tri, owner) <= 1.6 * G.edge_cut(tri, G.partition_rcb(tri.points, k))
            assert np.array_equal(owner, G.partition_graph(tri, k))  # deterministic
    # the cut of a 2-way split must be the edges between the parts (checked against a direct count)
    tri = G.triangulate_rectangle(0, 1, 0, 1, 12, 9, single_boundary=True)
    owner = G.partition_graph(tri, 2)
    T = tri.triangles
    e = np.unique(np.sort(np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]), axis=1), axis=0)
    assert G.edge_cut(tri, owner) == int((owner[e[:, 0]] != owner[e[:, 1]]).sum())
    with pytest.raises(G.FVMCudaError):
        G.partition_graph(tri, 0)


def test_patch_mesh_reproduces_du_at_its_nodes():
    """bench.py's in-run parity check of runs sharded by an arbitrary partition evaluates a patch (every triangle
    incident to the sampled nodes) as a mesh of its own: du at the sampled nodes must equal du of the global problem
    there -- checked here with the C oracle on an unstructured mesh and a graph partition's cut nodes."""
    from oracle.c_oracle import fvm_eqs_general
    from tests.common import delaunay_mesh
    tri = delaunay_mesh(3000, 4, jitter=0.3)
    owner = G.partition_graph(tri, 3)
    loc = G.extract_local(tri, owner, 1, 3)
    nodes = loc.global_nodes[np.unique(np.concatenate(loc.send_nodes))]
    patch, verts, pl = G.patch_mesh(tri, nodes)
    assert np.array_equal(verts[pl], nodes) and patch.num_triangles < tri.num_triangles // 2
    u = np.random.default_rng(3).random(tri.num_points)
    uvg, _ = tri.boundary_edges()
    dg = np.zeros(tri.num_points, bool)
    dg[np.unique(uvg)] = True
    full = fvm_eqs_general(tri.points, tri.triangles, u, flux_params=(0.7,), dirichlet=dg)
    uvp, _ = patch.boundary_edges()
    dp = np.zeros(patch.num_points, bool)
    dp[np.unique(uvp)] = True
    assert not dp[pl][~dg[nodes]].any()  # a sampled node keeps its complete triangle fan: it is interior to the patch
    part = fvm_eqs_general(patch.points, patch.triangles, u[verts], flux_params=(0.7,), dirichlet=dp)
    assert np.abs(part[pl] - full[nodes]).max() <= 1e-13 * np.abs(full).max()


def test_graph_partitioner_large_path_is_thread_count_independent():
    """Graphs above 131 072 nodes take the chunked parallel coarsening (fvm_partition.cu: coarsen_chunked); chunks are fixed
    by the node count, not by the thread count, so 1 thread and all threads must give the same owner array -- every rank of
    a sharded run computes the partition on its own slice of the cores and they must agree."""
    import subprocess
    import sys
    code = ("import sys, zlib, numpy as np; sys.path.insert(0, %r); import fvm_b200 as G; "
            "tri = G.triangulate_rectangle(0, 2, 0, 1, 520, 260, single_boundary=True); o = G.partition_graph(tri, 4); "
            "c = np.bincount(o, minlength=4); print(zlib.crc32(o.tobytes()), c.min(), c.max(), G.edge_cut(tri, o), "
            "G.edge_cut(tri, G.partition_rcb(tri.points, 4)))" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for nt in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.split())
    assert outs[0] == outs[1] == outs[2], outs
    crc, cmin, cmax, cut, cut_rcb = map(int, outs[0])
    assert cmax - cmin <= 1 and cut <= 1.3 * cut_rcb
