"""Triangulation container + the lattice generator used by the reference's examples.

DelaunayTriangulation.jl itself (unstructured generation, refinement) stays on the host side and
is out of scope; the engine accepts any flattened ccw triangle list.  `triangulate_rectangle`
restates DelaunayTriangulation.triangulate_rectangle."""
import numpy as np


class Triangulation:
    """points (N,2) f64, triangles (T,3) i32 (0-based, ccw, stored rotation kept),
    boundary_sections: list of ccw node sequences; section s has ghost vertex -(s+1)."""

    def __init__(self, points, triangles, boundary_sections=None, boundary_edge_list=None, num_sections=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
        if boundary_sections is None:
            boundary_sections = _chain_boundary(self.triangles)
        self.boundary_sections = [np.ascontiguousarray(s, dtype=np.int32) for s in boundary_sections]
        # rank-local meshes (sharding.py) carry an explicit edge list: only the part of the global
        # boundary that belongs to a local triangle, with the global section of every edge
        self._edge_list = None
        if boundary_edge_list is not None:
            uv, sec = boundary_edge_list
            self._edge_list = (np.ascontiguousarray(uv, dtype=np.int32).reshape(-1, 2), np.ascontiguousarray(sec, dtype=np.int32))
        self.num_sections = len(self.boundary_sections) if num_sections is None else int(num_sections)

    @property
    def num_points(self):
        return len(self.points)

    @property
    def num_triangles(self):
        return len(self.triangles)

    def boundary_edges(self):
        """keys(get_boundary_edge_map(tri)): (Eb,2) directed ccw edges and their section."""
        if self._edge_list is not None:
            return self._edge_list
        uv, sec = [], []
        for s, nodes in enumerate(self.boundary_sections):
            uv.append(np.stack([nodes[:-1], nodes[1:]], axis=1))
            sec.append(np.full(len(nodes) - 1, s, dtype=np.int32))
        if not uv:
            return np.zeros((0, 2), np.int32), np.zeros(0, np.int32)
        return np.ascontiguousarray(np.concatenate(uv), dtype=np.int32), np.concatenate(sec)

    def solid_vertex_mask(self):
        m = np.zeros(self.num_points, dtype=bool)
        m[self.triangles.ravel()] = True
        return m


def _chain_boundary(tris):
    """Directed edges without a reversed partner, chained into closed ccw loops."""
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]).astype(np.int64)
    n = int(e.max()) + 1
    key = e[:, 0] * n + e[:, 1]
    rkey = e[:, 1] * n + e[:, 0]
    bnd = e[~np.isin(key, rkey)]
    src = bnd[:, 0]
    if len(np.unique(src)) != len(src):
        # a vertex with two outgoing boundary edges (pinch point, or a hole touching the outer boundary): the loops are
        # ambiguous and DelaunayTriangulation would keep such a vertex in two sections -- pass boundary_sections explicitly
        raise ValueError("non-manifold boundary: a vertex has more than one outgoing boundary edge; give boundary_sections")
    nxt = dict(zip(src.tolist(), bnd[:, 1].tolist()))
    sections, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        cur = nxt[start]
        while cur != start:
            if cur in seen or cur not in nxt or len(loop) > len(bnd):
                raise ValueError("boundary edges do not form closed loops")
            loop.append(cur)
            seen.add(cur)
            cur = nxt[cur]
        loop.append(start)
        sections.append(loop)
    return sections


def triangulate_rectangle(a, b, c, d, nx, ny, single_boundary=False):
    dx = (b - a) / (nx - 1)
    dy = (d - c) / (ny - 1)
    pts = np.empty((nx * ny, 2))
    pts[:, 0] = np.tile(a + np.arange(nx, dtype=np.float64) * dx, ny)
    pts[:, 1] = np.repeat(c + np.arange(ny, dtype=np.float64) * dy, nx)
    p00 = (np.arange(nx - 1, dtype=np.int32)[None, :] + (np.arange(ny - 1, dtype=np.int32) * nx)[:, None]).ravel()
    tris = np.empty((2 * len(p00), 3), dtype=np.int32)
    tris[0::2, 0] = p00
    tris[0::2, 1] = p00 + 1
    tris[0::2, 2] = p00 + nx
    tris[1::2, 0] = p00 + nx
    tris[1::2, 1] = p00 + 1
    tris[1::2, 2] = p00 + nx + 1
    bottom = np.arange(0, nx)
    right = np.arange(nx - 1, nx * ny, nx)
    top = np.arange(nx * ny - 1, nx * (ny - 1) - 1, -1)
    left = np.arange(nx * (ny - 1), -1, -nx)
    if single_boundary:
        sections = [np.concatenate([bottom, right[1:], top[1:], left[1:]])]
    else:
        sections = [bottom, right, top, left]
    return Triangulation(pts, tris, sections)
