"""TEST INFRASTRUCTURE ONLY — independent restatement of the FVMWIRE container layout (include/fvmcuda.h,
"FVMWIRE" section) with struct + zlib, used by tests/ to check the C reader/writer byte for byte.
The reference (FiniteVolumeMethod.jl) has no file format; the arrays stored are the ones
`FVMGeometry(tri)` reads (/root/reference/src/geometry.jl:99-106) and `sol.u`, `sol.t`
(/root/reference/src/solve.jl:197-208).  Nothing in the product imports this module."""
import struct
import zlib

import numpy as np

MAGIC = b"FVMWIRE\0"
VERSION, ENDIAN, MAX_ARRAYS = 1, 0x01020304, 64
HEADER, ENTRY, ALIGN = 64, 96, 64
DATA_START = HEADER + ENTRY * MAX_ARRAYS
DTYPES = {1: np.dtype("<f8"), 2: np.dtype("<i4"), 3: np.dtype("u1"), 4: np.dtype("<i8")}
CODES = {v: k for k, v in DTYPES.items()}
_ENTRY_FMT = "<32sII4qQQII"
assert struct.calcsize(_ENTRY_FMT) == ENTRY and DATA_START % ALIGN == 0


def write(path, arrays):
    """arrays: list of (name, ndarray, dims-fastest-first or None)."""
    table, blobs, cursor = b"", [], DATA_START
    for name, a, dims in arrays:
        a = np.ascontiguousarray(a)
        dims = tuple(reversed(a.shape)) if dims is None else tuple(dims)
        raw = a.tobytes()  # little-endian host, like the C writer
        d4 = list(dims) + [0] * (4 - len(dims))
        table += struct.pack(_ENTRY_FMT, name.encode(), CODES[np.dtype(a.dtype)], len(dims), *d4, cursor, len(raw),
                             zlib.crc32(raw) & 0xFFFFFFFF, 0)
        pad = (-len(raw)) % ALIGN
        blobs.append(raw + b"\0" * pad)
        cursor += len(raw) + pad
    table += b"\0" * (ENTRY * MAX_ARRAYS - len(table))
    hdr = MAGIC + struct.pack("<IIIIQI", VERSION, ENDIAN, len(arrays), DATA_START, cursor, zlib.crc32(table) & 0xFFFFFFFF)
    hdr += b"\0" * (HEADER - len(hdr))
    with open(path, "wb") as f:
        f.write(hdr + table + b"".join(blobs))


def read(path):
    """-> dict name -> (ndarray shaped reversed(dims), dims).  Raises ValueError on any inconsistency."""
    raw = open(path, "rb").read()
    if len(raw) < DATA_START or raw[:8] != MAGIC:
        raise ValueError("not an FVMWIRE container")
    version, endian, n, start, nbytes, tcrc = struct.unpack("<IIIIQI", raw[8:36])
    if (version, endian, start) != (VERSION, ENDIAN, DATA_START) or n > MAX_ARRAYS or nbytes != len(raw):
        raise ValueError("corrupt header")
    if zlib.crc32(raw[HEADER:DATA_START]) & 0xFFFFFFFF != tcrc:
        raise ValueError("table checksum mismatch")
    out = {}
    for i in range(n):
        name, dt, rank, d0, d1, d2, d3, off, nb, crc, _ = struct.unpack(_ENTRY_FMT, raw[HEADER + i * ENTRY:HEADER + (i + 1) * ENTRY])
        dims = (d0, d1, d2, d3)[:rank]
        blob = raw[off:off + nb]
        if off % ALIGN or len(blob) != nb or zlib.crc32(blob) & 0xFFFFFFFF != crc:
            raise ValueError("array %r is corrupt" % name)
        out[name.rstrip(b"\0").decode()] = (np.frombuffer(blob, dtype=DTYPES[dt]).reshape(tuple(reversed(dims))), dims)
    return out
