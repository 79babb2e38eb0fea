"""FVMWIRE containers (include/fvmcuda.h, csrc/fvm_wire.cu): flat binary SoA files that move meshes
and solutions between the host application and libfvmcuda.

The reference has no file format of its own: a mesh is the DelaunayTriangulation object that
`FVMGeometry(tri)` walks (/root/reference/src/geometry.jl:99-106) and a solution is `sol.u`, `sol.t`
(/root/reference/src/solve.jl:197-208).  `save_mesh` / `load_mesh` and `save_solution` /
`load_solution` store exactly those arrays, column-major, 1-based like the Julia side writes them.
All byte handling is done by the C library; this module only marshals NumPy arrays."""
import ctypes as C

import numpy as np

from . import _lib as L
from .mesh import Triangulation

F64, I32, U8, I64 = 1, 2, 3, 4
_DTYPES = {F64: np.float64, I32: np.int32, U8: np.uint8, I64: np.int64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class WireError(L.FVMCudaError, OSError):
    """A container that is missing, truncated, corrupt or fails its checksum (FVM_ERR_IO)."""


def _check(w, rc):
    if rc != L.OK:
        msg = L.lib().fvm_wire_last_error(w).decode("utf-8", "replace")
        raise (WireError if rc == L.ERR_IO else L.FVMCudaError)(rc, msg)


class WireWriter:
    """Streams named arrays into a container; the header and table are written by close()."""

    def __init__(self, path):
        self._w = C.c_void_p()
        _check(None, L.lib().fvm_wire_create(str(path).encode(), C.byref(self._w)))

    def put(self, name, array, dims=None):
        """`dims` are given fastest-first (Julia's size(A)); by default a C-ordered NumPy array of shape
        (n, m) is stored with dims (m, n), i.e. as the column-major m-by-n matrix it is byte for byte."""
        a = np.ascontiguousarray(array)
        if a.dtype not in _CODES:
            raise TypeError("FVMWIRE stores float64, int32, uint8 and int64 arrays, not %s" % a.dtype)
        if dims is None:
            dims = tuple(reversed(a.shape))
        if int(np.prod(dims, dtype=np.int64)) != a.size:
            raise ValueError("dims %r do not match %d elements" % (dims, a.size))
        d = (C.c_int64 * 4)(*dims)
        _check(self._w, L.lib().fvm_wire_put(self._w, name.encode(), _CODES[a.dtype], len(dims), d, a.ctypes.data_as(C.c_void_p)))

    def close(self):
        if self._w:
            w, self._w = self._w, None
            _check(None, L.lib().fvm_wire_close(w))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class WireReader:
    def __init__(self, path):
        self._w = C.c_void_p()
        _check(None, L.lib().fvm_wire_open(str(path).encode(), C.byref(self._w)))
        n = C.c_int32()
        _check(self._w, L.lib().fvm_wire_count(self._w, C.byref(n)))
        self.arrays = {}
        for i in range(n.value):
            name = C.create_string_buffer(32)
            dt, rk, nb = C.c_int32(), C.c_int32(), C.c_int64()
            dims = (C.c_int64 * 4)()
            _check(self._w, L.lib().fvm_wire_info(self._w, i, name, C.byref(dt), C.byref(rk), dims, C.byref(nb)))
            self.arrays[name.value.decode()] = (i, dt.value, tuple(dims[:rk.value]), nb.value)

    def __contains__(self, name):
        return name in self.arrays

    def dims(self, name):
        return self.arrays[name][2]

    def get(self, name):
        """The array as a C-ordered NumPy array of shape reversed(dims) (see WireWriter.put)."""
        if name not in self.arrays:
            idx = C.c_int32()
            _check(self._w, L.lib().fvm_wire_find(self._w, name.encode(), C.byref(idx)))
        i, dt, dims, nb = self.arrays[name]
        out = np.empty(tuple(reversed(dims)), dtype=_DTYPES[dt])
        _check(self._w, L.lib().fvm_wire_get(self._w, i, out.ctypes.data_as(C.c_void_p), nb))
        return out

    def close(self):
        if self._w:
            w, self._w = self._w, None
            _check(None, L.lib().fvm_wire_close(w))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save_mesh(path, tri, index_base=1):
    """points (2,N), triangles (3,T), boundary sections and boundary edges of a Triangulation."""
    uv, sec = tri.boundary_edges()
    ptr = np.zeros(len(tri.boundary_sections) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(s) for s in tri.boundary_sections])
    nodes = np.concatenate(tri.boundary_sections).astype(np.int32) if tri.boundary_sections else np.zeros(0, np.int32)
    with WireWriter(path) as w:
        w.put("points", tri.points)
        w.put("triangles", tri.triangles + np.int32(index_base))
        w.put("index_base", np.array([index_base], dtype=np.int32))
        w.put("boundary_ptr", ptr)
        w.put("boundary_nodes", nodes + np.int32(index_base))
        w.put("boundary_edges", uv + np.int32(index_base))
        w.put("boundary_edge_section", sec.astype(np.int32))
        w.put("num_sections", np.array([tri.num_sections], dtype=np.int32))


def load_mesh(path):
    with WireReader(path) as r:
        base = int(r.get("index_base")[0]) if "index_base" in r else 1
        pts = r.get("points")
        tris = r.get("triangles") - np.int32(base)
        sections = None
        if "boundary_ptr" in r:
            ptr, nodes = r.get("boundary_ptr"), r.get("boundary_nodes") - np.int32(base)
            sections = [nodes[ptr[s]:ptr[s + 1]] for s in range(len(ptr) - 1)]
        edge_list = None
        if "boundary_edges" in r and "boundary_edge_section" in r:
            edge_list = (r.get("boundary_edges") - np.int32(base), r.get("boundary_edge_section"))
        nsec = int(r.get("num_sections")[0]) if "num_sections" in r else None
    if sections is not None and edge_list is not None:
        # a self-contained mesh derives its edges from the sections; keep an explicit list only when it
        # differs (rank-local meshes of sharding.py)
        t = Triangulation(pts, tris, sections, num_sections=nsec)
        uv, sec = t.boundary_edges()
        if uv.shape == edge_list[0].shape and np.array_equal(uv, edge_list[0]) and np.array_equal(sec, edge_list[1]):
            return t
    return Triangulation(pts, tris, sections, boundary_edge_list=edge_list, num_sections=nsec)


def save_solution(path, sol, neq=1):
    """sol.u: (nsave, N) / (nsave, N*neq) states or a single vector; sol.t: times (optional)."""
    u = np.asarray(sol.u, dtype=np.float64)
    with WireWriter(path) as w:
        if u.ndim == 1:
            w.put("u", u, dims=(u.size // neq, 1) if neq == 1 else (neq, u.size // neq, 1))
        else:
            n = u.shape[1] // neq
            w.put("u", u, dims=(n, u.shape[0]) if neq == 1 else (neq, n, u.shape[0]))
        if getattr(sol, "t", None) is not None:
            w.put("t", np.atleast_1d(np.asarray(sol.t, dtype=np.float64)))
        for key, code in (("iters", np.int64), ("relres", np.float64)):
            v = getattr(sol, key, None)
            if v is not None:
                w.put(key, np.array([v], dtype=code))
        w.put("retcode", np.frombuffer(str(getattr(sol, "retcode", "Success")).encode(), dtype=np.uint8))


def load_solution(path):
    from .templates import Solution
    with WireReader(path) as r:
        u = r.get("u")
        u = u.reshape(u.shape[0], -1)
        t = r.get("t") if "t" in r else None
        iters = int(r.get("iters")[0]) if "iters" in r else None
        relres = float(r.get("relres")[0]) if "relres" in r else None
        retcode = r.get("retcode").tobytes().decode() if "retcode" in r else "Success"
    return Solution(u, t, iters, relres, retcode)
