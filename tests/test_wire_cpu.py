"""FVMWIRE container: the C reader/writer of libfvmcuda against the independent
struct+zlib restatement in oracle/fvm_wire.py — byte-exact files, checksums, corruption and truncation
handling, mesh / solution round trips.  Host-only: needs no GPU."""
import os
import zlib

import numpy as np
import pytest

import fvm_b200 as G
import fvm_b200._lib as L
from oracle import fvm_wire as OW
from tests.common import delaunay_mesh


def _sample_arrays(rng):
    return [
        ("points", rng.random((37, 2)), None),
        ("triangles", rng.integers(1, 38, size=(51, 3)).astype(np.int32), None),
        ("kind", rng.integers(0, 3, size=101).astype(np.uint8), None),
        ("offsets", np.arange(7, dtype=np.int64) * 2**33, None),
        ("empty", np.zeros(0, dtype=np.float64), None),
        ("u", rng.random((3, 37 * 2)), (2, 37, 3)),
        ("scalar", np.array([1], dtype=np.int32), None),
    ]


def test_crc32_is_zlib_crc32():
    rng = np.random.default_rng(1)
    for n in list(range(0, 40)) + [255, 256, 257, 4099, 1 << 16]:
        buf = rng.integers(0, 256, size=n).astype(np.uint8)
        got = L.lib().fvm_wire_crc32(buf.ctypes.data, n)
        assert got == (zlib.crc32(buf.tobytes()) & 0xFFFFFFFF), n
    assert L.lib().fvm_wire_crc32(b"123456789", 9) == 0xCBF43926  # the CRC-32/ISO-HDLC check value


def test_c_writer_is_byte_identical_to_the_restatement(tmp_path):
    arrays = _sample_arrays(np.random.default_rng(2))
    pc, po = str(tmp_path / "c.fvmw"), str(tmp_path / "o.fvmw")
    with G.WireWriter(pc) as w:
        for name, a, dims in arrays:
            w.put(name, a, dims)
    OW.write(po, arrays)
    bc, bo = open(pc, "rb").read(), open(po, "rb").read()
    assert len(bc) == len(bo) and bc == bo
    assert len(bc) % 64 == 0


def test_c_reader_reads_the_restatement_and_back(tmp_path):
    arrays = _sample_arrays(np.random.default_rng(3))
    po = str(tmp_path / "o.fvmw")
    OW.write(po, arrays)
    with G.WireReader(po) as r:
        assert list(r.arrays) == [a[0] for a in arrays]
        for name, a, dims in arrays:
            got = r.get(name)
            want_dims = tuple(reversed(a.shape)) if dims is None else dims
            assert r.dims(name) == want_dims
            assert got.dtype == a.dtype and np.array_equal(got.ravel(), a.ravel())
    pc = str(tmp_path / "c.fvmw")
    with G.WireWriter(pc) as w:
        for name, a, dims in arrays:
            w.put(name, a, dims)
    back = OW.read(pc)
    for name, a, dims in arrays:
        assert np.array_equal(back[name][0].ravel(), a.ravel())


def test_payloads_are_64_byte_aligned(tmp_path):
    p = str(tmp_path / "a.fvmw")
    arrays = _sample_arrays(np.random.default_rng(4))
    with G.WireWriter(p) as w:
        for name, a, dims in arrays:
            w.put(name, a, dims)
    raw = open(p, "rb").read()
    import struct
    n = struct.unpack("<I", raw[16:20])[0]
    assert n == len(arrays)
    for i in range(n):
        off, nb = struct.unpack("<QQ", raw[64 + i * 96 + 72:64 + i * 96 + 88])
        assert off % 64 == 0 and off >= OW.DATA_START and off + nb <= len(raw)


def test_corruption_and_truncation_are_detected(tmp_path):
    p = str(tmp_path / "a.fvmw")
    arrays = _sample_arrays(np.random.default_rng(5))
    OW.write(p, arrays)
    raw = bytearray(open(p, "rb").read())
    # one flipped payload bit: open succeeds, get of that array fails with the checksum error
    bad = bytearray(raw)
    bad[OW.DATA_START + 5] ^= 0x10
    q = str(tmp_path / "payload.fvmw")
    open(q, "wb").write(bad)
    with G.WireReader(q) as r:
        with pytest.raises(G.WireError, match="checksum mismatch in array points"):
            r.get("points")
        assert np.array_equal(r.get("kind"), arrays[2][1])  # other arrays are still readable
    # flipped table byte
    bad = bytearray(raw)
    bad[64 + 40] ^= 0x01
    open(q, "wb").write(bad)
    with pytest.raises(G.WireError, match="table checksum"):
        G.WireReader(q)
    # truncated file
    open(q, "wb").write(raw[:-64])
    with pytest.raises(G.WireError, match="truncated"):
        G.WireReader(q)
    open(q, "wb").write(raw[:100])
    with pytest.raises(G.WireError, match="shorter"):
        G.WireReader(q)
    # foreign magic / byte order / version
    bad = bytearray(raw)
    bad[0:4] = b"NOPE"
    open(q, "wb").write(bad)
    with pytest.raises(G.WireError, match="not an FVMWIRE"):
        G.WireReader(q)
    bad = bytearray(raw)
    bad[12:16] = bad[12:16][::-1]
    open(q, "wb").write(bad)
    with pytest.raises(G.WireError, match="byte order"):
        G.WireReader(q)
    with pytest.raises(G.WireError, match="cannot open"):
        G.WireReader(str(tmp_path / "missing.fvmw"))


def test_writer_argument_errors(tmp_path):
    p = str(tmp_path / "a.fvmw")
    with G.WireWriter(p) as w:
        w.put("x", np.zeros(3))
        with pytest.raises(G.FVMCudaError, match="duplicate"):
            w.put("x", np.zeros(3))
        with pytest.raises(G.FVMCudaError, match="name must"):
            w.put("n" * 32, np.zeros(3))
        with pytest.raises(TypeError):
            w.put("f32", np.zeros(3, dtype=np.float32))
        with pytest.raises(ValueError):
            w.put("y", np.zeros(6), dims=(4, 2))
        for i in range(63):
            w.put("a%d" % i, np.zeros(1))
        with pytest.raises(G.FVMCudaError, match="table is full"):
            w.put("overflow", np.zeros(1))
    with G.WireReader(p) as r:
        assert len(r.arrays) == 64
        with pytest.raises(G.FVMCudaError, match="no array named"):
            r.get("absent")


@pytest.mark.parametrize("kind", ["lattice1", "lattice4", "delaunay"])
def test_mesh_round_trip_is_bit_exact(tmp_path, kind):
    if kind == "lattice1":
        tri = G.triangulate_rectangle(0.0, 2.0, -1.0, 3.0, 12, 19, single_boundary=True)
    elif kind == "lattice4":
        tri = G.triangulate_rectangle(0.0, 2.0, -1.0, 3.0, 12, 19, single_boundary=False)
    else:
        tri = delaunay_mesh(300, seed=7)
    p = str(tmp_path / "mesh.fvmw")
    G.save_mesh(p, tri)
    back = G.load_mesh(p)
    assert np.array_equal(back.points, tri.points) and back.points.dtype == np.float64
    assert np.array_equal(back.triangles, tri.triangles) and back.triangles.dtype == np.int32
    assert len(back.boundary_sections) == len(tri.boundary_sections) and back.num_sections == tri.num_sections
    for a, b in zip(back.boundary_sections, tri.boundary_sections):
        assert np.array_equal(a, b)
    for a, b in zip(back.boundary_edges(), tri.boundary_edges()):
        assert np.array_equal(a, b)
    # on disk the indices are 1-based (what the Julia side writes), column-major (3,T) / (2,N)
    onfile = OW.read(p)
    assert onfile["triangles"][1] == (3, tri.num_triangles) and onfile["points"][1] == (2, tri.num_points)
    assert onfile["triangles"][0].min() == 1 and int(onfile["index_base"][0][0]) == 1


def test_rank_local_mesh_round_trip(tmp_path):
    """A shard's local mesh carries an explicit boundary-edge list with global section ids."""
    tri = G.triangulate_rectangle(0.0, 1.0, 0.0, 1.0, 9, 13, single_boundary=False)
    owner = G.partition_strips(tri.points, 2)
    lm = G.extract_local(tri, owner, 1)
    p = str(tmp_path / "local.fvmw")
    G.save_mesh(p, lm.triangulation)
    back = G.load_mesh(p)
    assert np.array_equal(back.points, lm.triangulation.points) and np.array_equal(back.triangles, lm.triangulation.triangles)
    for a, b in zip(back.boundary_edges(), lm.triangulation.boundary_edges()):
        assert np.array_equal(a, b)
    assert back.num_sections == lm.triangulation.num_sections


def test_solution_round_trip(tmp_path):
    rng = np.random.default_rng(11)
    p = str(tmp_path / "sol.fvmw")
    sol = G.Solution(rng.random((5, 40)), t=np.linspace(0, 1, 5))
    G.save_solution(p, sol)
    back = G.load_solution(p)
    assert np.array_equal(back.u, sol.u) and np.array_equal(back.t, sol.t) and back.retcode == "Success"
    assert OW.read(p)["u"][1] == (40, 5)
    sysol = G.Solution(rng.random((3, 2 * 40)), t=np.array([0.0, 0.5, 1.0]))
    G.save_solution(p, sysol, neq=2)
    assert OW.read(p)["u"][1] == (2, 40, 3)
    assert np.array_equal(G.load_solution(p).u, sysol.u)
    steady = G.Solution(rng.random(40), iters=17, relres=1e-11, retcode="Success")
    G.save_solution(p, steady)
    back = G.load_solution(p)
    assert np.array_equal(back.u[0], steady.u) and back.iters == 17 and back.relres == 1e-11 and back.t is None


def test_create_from_wire_reports_errors_without_touching_the_gpu(tmp_path):
    import ctypes as C
    h = C.c_void_p()
    rc = L.lib().fvm_create_from_wire(str(tmp_path / "missing.fvmw").encode(), 1, 0, C.byref(h))
    assert rc == L.ERR_IO and not h.value
    assert b"cannot open" in L.lib().fvm_wire_last_error(None)
    p = str(tmp_path / "nomesh.fvmw")
    with G.WireWriter(p) as w:
        w.put("points", np.zeros((4, 3)))  # wrong leading dimension
        w.put("triangles", np.zeros((2, 3), dtype=np.int32))
    rc = L.lib().fvm_create_from_wire(p.encode(), 1, 0, C.byref(h))
    assert rc == L.ERR_ARG and b"wrong type or shape" in L.lib().fvm_wire_last_error(None)
