"""Multi-GPU sharding of a problem: node partition, local meshes with a one-layer ghost halo, and
the halo-exchange plan.  One process per GPU; the device side is in csrc/fvm_shard.cu.

Every rank owns a set of nodes and keeps ALL triangles that touch an owned node, so the right-hand
side of an owned node is complete locally and only the input vector's ghost entries are exchanged.
The reference has no distributed code at all; this layer is new (north_star d)."""
import ctypes as C

import numpy as np

from . import _lib as L
from .conditions import BoundaryConditions, InternalConditions
from .mesh import Triangulation
from .problem import FVMGeometry, FVMProblem, FVMSystem, get_cuda_parameters


# ---- partitioners --------------------------------------------------------------------------------
def partition_strips(points, nparts, axis=1):
    """Equal-count strips along one axis (row strips of a lattice when nparts divides ny)."""
    n = len(points)
    order = np.lexsort((points[:, 1 - axis], points[:, axis]))
    owner = np.empty(n, dtype=np.int32)
    owner[order] = (np.arange(n, dtype=np.int64) * nparts // n).astype(np.int32)
    return owner


def partition_rcb(points, nparts):
    """Recursive coordinate bisection (the METIS-style graph partitioner named by north_star is not
    available offline; RCB gives compact subdomains with short interfaces on these 2-D meshes)."""
    owner = np.zeros(len(points), dtype=np.int32)

    def rec(idx, lo, hi):
        if hi - lo == 1:
            owner[idx] = lo
            return
        mid = (lo + hi) // 2
        p = points[idx]
        axis = int(np.argmax(p.max(axis=0) - p.min(axis=0)))
        k = len(idx) * (mid - lo) // (hi - lo)
        part = np.argpartition(p[:, axis], k - 1 if k > 0 else 0)
        rec(idx[part[:k]], lo, mid)
        rec(idx[part[k:]], mid, hi)

    rec(np.arange(len(points)), 0, nparts)
    return owner


def partition_graph(tri, nparts):
    """METIS-style partition of the node graph (multilevel recursive bisection in the C library: heavy-edge
    matching, greedy graph growing, FM refinement): parts differ by at most one node; the edge cut — and with
    it the halo — follows the mesh connectivity instead of the coordinates."""
    owner = np.empty(tri.num_points, dtype=np.int32)
    t = L.i32(tri.triangles)
    rc = L.lib().fvm_partition_graph(tri.num_points, L.ip(t), tri.num_triangles, 0, int(nparts), L.ip(owner))
    if rc != L.OK:
        raise L.FVMCudaError(rc, L.lib().fvm_last_error(None).decode())
    return owner


def edge_cut(tri, owner):
    """number of node-graph edges between different parts"""
    cut = C.c_int64()
    t, o = L.i32(tri.triangles), L.i32(owner)
    rc = L.lib().fvm_partition_edge_cut(tri.num_points, L.ip(t), tri.num_triangles, 0, L.ip(o), C.byref(cut))
    if rc != L.OK:
        raise L.FVMCudaError(rc, L.lib().fvm_last_error(None).decode())
    return cut.value


# ---- local mesh + halo plan ------------------------------------------------------------------------
class LocalMesh:
    """rank-local view: `global_nodes` (owned first, then ghosts, both ascending global id),
    `triangles` in local ids, boundary edges with their global section, and the halo plan:
    `neighbours[q]`, `send_nodes[q]` / `recv_nodes[q]` (local ids, ascending global id)."""

    def __init__(self, rank, nparts, points, triangles, global_nodes, n_owned, global_triangles, bedges, bsec, num_sections,
                 neighbours, send_nodes, recv_nodes):
        self.rank, self.nparts = rank, nparts
        self.global_nodes, self.n_owned, self.global_triangles = global_nodes, n_owned, global_triangles
        self.is_ghost = np.zeros(len(global_nodes), dtype=np.uint8)
        self.is_ghost[n_owned:] = 1
        self.neighbours, self.send_nodes, self.recv_nodes = neighbours, send_nodes, recv_nodes
        self.triangulation = Triangulation(points, triangles, boundary_sections=[], boundary_edge_list=(bedges, bsec),
                                           num_sections=num_sections)

    def halo_arrays(self):
        sp = np.zeros(len(self.neighbours) + 1, dtype=np.int32)
        rp = np.zeros(len(self.neighbours) + 1, dtype=np.int32)
        for q in range(len(self.neighbours)):
            sp[q + 1] = sp[q] + len(self.send_nodes[q])
            rp[q + 1] = rp[q] + len(self.recv_nodes[q])
        sn = np.concatenate(self.send_nodes).astype(np.int32) if self.neighbours else np.zeros(0, np.int32)
        rn = np.concatenate(self.recv_nodes).astype(np.int32) if self.neighbours else np.zeros(0, np.int32)
        return np.asarray(self.neighbours, dtype=np.int32), sp, sn, rp, rn


def extract_local(tri, owner, rank, nparts=None):
    """Local mesh of `rank` from the global triangulation and a node -> rank map."""
    owner = np.asarray(owner)
    nparts = int(owner.max()) + 1 if nparts is None else nparts
    T = tri.triangles
    N = tri.num_points
    own_t = owner[T]  # (T,3)
    mine_t = (own_t == rank).any(axis=1)
    gt = np.nonzero(mine_t)[0]
    lt = T[gt]
    verts = np.unique(lt)
    owned = np.nonzero(owner == rank)[0]  # includes owned points that are not vertices
    ghosts = verts[owner[verts] != rank]
    gnodes = np.concatenate([owned, ghosts]).astype(np.int64)
    g2l = np.full(N, -1, dtype=np.int64)
    g2l[gnodes] = np.arange(len(gnodes))
    ltri = g2l[lt].astype(np.int32)
    # boundary edges of the global mesh that belong to a local triangle
    uv, sec = tri.boundary_edges()
    if len(uv):
        keys = uv[:, 0].astype(np.int64) * N + uv[:, 1]
        srt = np.argsort(keys)
        de = np.concatenate([lt[:, [0, 1]], lt[:, [1, 2]], lt[:, [2, 0]]]).astype(np.int64)
        dk = de[:, 0] * N + de[:, 1]
        pos = np.searchsorted(keys[srt], dk)
        pos[pos >= len(keys)] = 0
        hit = keys[srt][pos] == dk
        sel = np.sort(srt[pos[hit]])  # keep the global edge order
        bed = g2l[uv[sel]].astype(np.int32)
        bsec = sec[sel]
    else:
        bed, bsec = np.zeros((0, 2), np.int32), np.zeros(0, np.int32)
    # halo plan: vertex v is a ghost on rank owner[w] for every triangle mate w with another owner
    cut = (own_t != own_t[:, [0]]).any(axis=1)
    ct, co = T[cut], own_t[cut]
    pairs = []
    for a in range(3):
        for b in range(3):
            if a != b:
                m = co[:, a] != co[:, b]
                pairs.append(np.stack([ct[m, a], co[m, b]], axis=1))
    pairs = np.unique(np.concatenate(pairs), axis=0) if pairs and len(ct) else np.zeros((0, 2), np.int64)
    neighbours, send_nodes, recv_nodes = [], [], []
    for q in range(nparts):
        if q == rank:
            continue
        snd = pairs[(pairs[:, 1] == q) & (owner[pairs[:, 0]] == rank), 0]  # my nodes that are ghosts on q
        rcv = pairs[(pairs[:, 1] == rank) & (owner[pairs[:, 0]] == q), 0]  # q's nodes that are ghosts here
        if len(snd) or len(rcv):
            neighbours.append(q)
            send_nodes.append(g2l[np.sort(snd)])
            recv_nodes.append(g2l[np.sort(rcv)])
    return LocalMesh(rank, nparts, tri.points[gnodes], ltri, gnodes, len(owned), gt, bed, bsec, len(tri.boundary_sections),
                     neighbours, send_nodes, recv_nodes)


def lattice_strip_local(a, b, c, d, nx, ny_per_rank, rank, nparts):
    """Local mesh of a rank of the weak-scaling lattice [a,b]x[c,d] with nx x (ny_per_rank*nparts)
    nodes split into row strips, built without materialising the global mesh (bench.py, config 5)."""
    ny = ny_per_rank * nparts
    j0, j1 = rank * ny_per_rank, (rank + 1) * ny_per_rank  # owned rows
    g0, g1 = max(0, j0 - 1), min(ny, j1 + 1)               # with ghost rows
    dy = (d - c) / (ny - 1)
    dx = (b - a) / (nx - 1)
    rows = np.arange(g0, g1)
    owned_rows = (rows >= j0) & (rows < j1)
    order = np.concatenate([rows[owned_rows], rows[~owned_rows]])  # owned first, then ghosts (ascending)
    nloc_rows = len(order)
    row_pos = np.empty(ny, dtype=np.int64)
    row_pos[order] = np.arange(nloc_rows)
    pts = np.empty((nloc_rows * nx, 2))
    pts[:, 0] = np.tile(a + np.arange(nx, dtype=np.float64) * dx, nloc_rows)
    pts[:, 1] = np.repeat(c + order.astype(np.float64) * dy, nx)
    gnodes = (order[:, None].astype(np.int64) * nx + np.arange(nx)[None, :]).ravel()
    jj = np.arange(g0, g1 - 1)  # cell rows
    p00 = (row_pos[jj][:, None] * nx + np.arange(nx - 1)[None, :]).ravel()
    p01 = (row_pos[jj + 1][:, None] * nx + np.arange(nx - 1)[None, :]).ravel()
    tris = np.empty((2 * len(p00), 3), dtype=np.int32)
    tris[0::2, 0], tris[0::2, 1], tris[0::2, 2] = p00, p00 + 1, p01
    tris[1::2, 0], tris[1::2, 1], tris[1::2, 2] = p01, p00 + 1, p01 + 1
    # boundary edges (single section, ccw): bottom / top only on the first / last rank, sides on all
    ed = []
    if rank == 0:
        i = np.arange(nx - 1)
        ed.append(np.stack([row_pos[0] * nx + i, row_pos[0] * nx + i + 1], axis=1))
    ed.append(np.stack([row_pos[jj] * nx + nx - 1, row_pos[jj + 1] * nx + nx - 1], axis=1))  # right side, upwards
    if rank == nparts - 1:
        i = np.arange(nx - 1, 0, -1)
        ed.append(np.stack([row_pos[ny - 1] * nx + i, row_pos[ny - 1] * nx + i - 1], axis=1))
    ed.append(np.stack([row_pos[jj[::-1] + 1] * nx, row_pos[jj[::-1]] * nx], axis=1))  # left side, downwards
    bed = np.concatenate(ed).astype(np.int32)
    neighbours, send_nodes, recv_nodes = [], [], []
    for q, (srow, rrow) in ((rank - 1, (j0, j0 - 1)), (rank + 1, (j1 - 1, j1))):
        if 0 <= q < nparts:
            neighbours.append(q)
            send_nodes.append(row_pos[srow] * nx + np.arange(nx))
            recv_nodes.append(row_pos[rrow] * nx + np.arange(nx))
    n_owned = int(owned_rows.sum()) * nx
    gt = None
    return LocalMesh(rank, nparts, pts, tris, gnodes, n_owned, gt, bed, np.zeros(len(bed), np.int32), 1, neighbours, send_nodes,
                     recv_nodes)


def lattice_rows(a, b, c, d, nx, ny, row0, row1):
    """Rows [row0, row1) of the nx x ny lattice on [a,b]x[c,d] as a mesh of their own, with the GLOBAL coordinates
    (same doubles as triangulate_rectangle) and the same triangle pattern: du of an inner row of the patch equals du
    of that row in the global problem (bench.py's in-run parity check of sharded runs).  One boundary section."""
    dx = (b - a) / (nx - 1)
    dy = (d - c) / (ny - 1)
    rows = np.arange(row0, row1)
    nr = len(rows)
    pts = np.empty((nr * nx, 2))
    pts[:, 0] = np.tile(a + np.arange(nx, dtype=np.float64) * dx, nr)
    pts[:, 1] = np.repeat(c + rows.astype(np.float64) * dy, nx)
    p00 = (np.arange(nx - 1, dtype=np.int32)[None, :] + (np.arange(nr - 1, dtype=np.int32) * nx)[:, None]).ravel()
    tris = np.empty((2 * len(p00), 3), dtype=np.int32)
    tris[0::2, 0], tris[0::2, 1], tris[0::2, 2] = p00, p00 + 1, p00 + nx
    tris[1::2, 0], tris[1::2, 1], tris[1::2, 2] = p00 + nx, p00 + 1, p00 + nx + 1
    bottom = np.arange(0, nx)
    right = np.arange(nx - 1, nx * nr, nx)
    top = np.arange(nx * nr - 1, nx * (nr - 1) - 1, -1)
    left = np.arange(nx * (nr - 1), -1, -nx)
    return Triangulation(pts, tris, [np.concatenate([bottom, right[1:], top[1:], left[1:]])])


def patch_mesh(tri, nodes):
    """The sub-mesh of every triangle incident to one of `nodes` (global ids), as a mesh of its own with the global
    coordinates: `nodes` keep their complete triangle fans, so du of the global problem at those nodes equals du of the
    patch problem there (the patch's rim is a Dirichlet boundary whose values are ignored).  Used by bench.py's in-run
    parity check of runs sharded by an arbitrary partition.  Returns (Triangulation, patch -> global ids, local ids of
    `nodes`)."""
    N = tri.num_points
    nodes = np.asarray(nodes, dtype=np.int64)
    mark = np.zeros(N, dtype=bool)
    mark[nodes] = True
    lt = tri.triangles[mark[tri.triangles].any(axis=1)]
    verts = np.unique(lt)
    g2l = np.full(N, -1, dtype=np.int32)
    g2l[verts] = np.arange(len(verts), dtype=np.int32)
    ltri = g2l[lt]
    e = np.concatenate([ltri[:, [0, 1]], ltri[:, [1, 2]], ltri[:, [2, 0]]]).astype(np.int64)
    n = len(verts)
    bnd = e[~np.isin(e[:, 0] * n + e[:, 1], e[:, 1] * n + e[:, 0])]  # directed edges without a reversed partner
    patch = Triangulation(tri.points[verts], ltri, boundary_sections=[], boundary_edge_list=(bnd, np.zeros(len(bnd), np.int32)),
                          num_sections=1)
    return patch, verts, g2l[nodes]


# ---- problems ---------------------------------------------------------------------------------------
def shard_problem(prob, local):
    """The rank-local FVMProblem / FVMSystem of a global problem (same condition functions, sections
    and functors; initial condition restricted to the local nodes)."""
    mesh = FVMGeometry(local.triangulation)
    g2l = {int(g): i for i, g in enumerate(local.global_nodes)}

    def one(p):
        c = p.conditions
        nif = c.nif
        bc = BoundaryConditions(mesh, tuple(c.functions[nif:]), c.condition_types)
        ic = InternalConditions(tuple(c.functions[:nif]),
                                dirichlet_nodes={g2l[n]: f for n, f in c.internal.dirichlet_nodes.items() if n in g2l},
                                dudt_nodes={g2l[n]: f for n, f in c.internal.dudt_nodes.items() if n in g2l})
        return FVMProblem(mesh, bc, ic, flux_function=p.flux_function, source_function=p.source_function,
                          initial_condition=p.initial_condition[local.global_nodes], initial_time=p.initial_time,
                          final_time=p.final_time)

    probs = [one(p) for p in prob.problems]
    return probs[0] if prob.neqs == 0 else FVMSystem(*probs)


def get_sharded_cuda_parameters(local_prob, local, dist=None, **kw):
    """get_cuda_parameters for a rank-local problem: flags the ghost layer, joins the NCCL
    communicator (unique id broadcast through torch.distributed) and installs the halo plan."""
    p = get_cuda_parameters(local_prob, ghost=local.is_ghost, **kw)
    install_halo(p.engine, local, dist)
    return p


def install_halo(engine, local, dist=None):
    lib = L.lib()
    if local.nparts > 1:
        import ctypes as C
        if dist is None:
            import torch.distributed as dist
        # a problem one rank rejects (its local conditions hold e.g. a Dudt node a steady template does not support) is
        # rejected by EVERY rank, before anyone enters a collective
        verdicts = [None] * local.nparts
        dist.all_gather_object(verdicts, getattr(engine, "validation_error", None))
        bad = [(r, v) for r, v in enumerate(verdicts) if v]
        if bad:
            engine.close()
            raise ValueError(bad[0][1] + (" (conditions of rank %d)" % bad[0][0] if bad[0][0] != local.rank else ""))
        uid = C.create_string_buffer(128)
        if local.rank == 0:
            rc = lib.fvm_nccl_unique_id(uid)
            if rc != L.OK:
                raise L.FVMCudaError(rc, "ncclGetUniqueId failed")
        box = [uid.raw]
        dist.broadcast_object_list(box, src=0)
        uid = C.create_string_buffer(box[0], 128)
        L.check(engine.h, lib.fvm_shard_init(engine.h, uid, local.rank, local.nparts))
    nb, sp, sn, rp, rn = local.halo_arrays()
    L.check(engine.h, lib.fvm_set_halo(engine.h, len(nb), L.ip(nb), L.ip(sp), L.ip(sn), L.ip(rp), L.ip(rn)))
    engine._keep.append((nb, sp, sn, rp, rn))
