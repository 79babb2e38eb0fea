// libfvmcuda: the FVMWIRE container — flat binary SoA arrays with a small header — that moves
// 10^7..10^8-node meshes and solutions between the Julia host and the library.
// The reference has no on-disk format of its own: a mesh is a DelaunayTriangulation object
// (/root/reference/src/geometry.jl:99-106 reads it through each_point / each_solid_triangle) and a
// solution is `sol.u::Vector{Vector{Float64}}` (/root/reference/src/solve.jl:197-208); the arrays
// stored here are exactly those, column-major, in the caller's numbering.
//
// Layout (all integers little-endian as written by the host; the reader rejects a foreign endian tag):
//   [0,64)    header : magic "FVMWIRE\0", u32 version, u32 endian tag 0x01020304, u32 n_arrays,
//                      u32 data_start, u64 file_bytes, u32 crc32(table), 28 reserved zero bytes
//   [64, ...) table  : FVM_WIRE_MAX_ARRAYS entries of 96 bytes (unused entries are zero)
//                      char name[32], u32 dtype, u32 rank, i64 dims[4] (dims[0] fastest = Julia order),
//                      u64 offset, u64 nbytes, u32 crc32(data), u32 reserved
//   data_start       : array payloads, each at a 64-byte aligned offset, zero padding between them
// Host-only code: no kernel here, and nothing on the compute path reads or writes files.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "fvm_internal.h"

namespace {

constexpr uint32_t WIRE_VERSION = 1;
constexpr uint32_t WIRE_ENDIAN = 0x01020304u;
constexpr size_t WIRE_HEADER = 64, WIRE_ENTRY = 96, WIRE_ALIGN = 64;
constexpr size_t WIRE_DATA_START = WIRE_HEADER + WIRE_ENTRY * FVM_WIRE_MAX_ARRAYS;  // 6208 = 97 * 64
const char WIRE_MAGIC[8] = {'F', 'V', 'M', 'W', 'I', 'R', 'E', '\0'};

struct WireEntry {
    char name[32];
    uint32_t dtype, rank;
    int64_t dims[4];
    uint64_t offset, nbytes;
    uint32_t crc, reserved;
};
static_assert(sizeof(WireEntry) == WIRE_ENTRY, "table entry layout");
static_assert(WIRE_DATA_START % WIRE_ALIGN == 0, "payload alignment");

// CRC-32 (IEEE 802.3, reflected, poly 0xEDB88320 — zlib's crc32), slicing-by-8
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    }
};
const CrcTables& crc_tables() {
    static const CrcTables T;
    return T;
}
uint32_t crc32_update(uint32_t crc, const void* data, size_t n) {
    const CrcTables& T = crc_tables();
    const uint8_t* p = static_cast<const uint8_t*>(data);
    crc = ~crc;
    while (n >= 8) {
        uint32_t a, b;
        memcpy(&a, p, 4);
        memcpy(&b, p + 4, 4);
        a ^= crc;
        crc = T.t[7][a & 0xFF] ^ T.t[6][(a >> 8) & 0xFF] ^ T.t[5][(a >> 16) & 0xFF] ^ T.t[4][a >> 24] ^ T.t[3][b & 0xFF] ^
              T.t[2][(b >> 8) & 0xFF] ^ T.t[1][(b >> 16) & 0xFF] ^ T.t[0][b >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) crc = T.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

size_t dtype_size(int32_t dtype) {
    switch (dtype) {
        case FVM_WIRE_F64: return 8;
        case FVM_WIRE_I32: return 4;
        case FVM_WIRE_U8: return 1;
        case FVM_WIRE_I64: return 8;
    }
    return 0;
}

}  // namespace

struct fvm_wire {
    FILE* f = nullptr;
    bool writing = false;
    std::string path, err;
    std::vector<WireEntry> table;
    uint64_t cursor = 0;  // writer: next free byte
    uint64_t file_bytes = 0;
};

static thread_local std::string g_wire_err;  // errors of calls that have no handle yet

static int32_t wire_fail(fvm_wire* w, int32_t code, const std::string& msg) {
    if (w) w->err = msg;
    g_wire_err = msg;
    return code;
}

static void pack_header(const fvm_wire* w, uint8_t* hdr /*64*/, const std::vector<uint8_t>& table_bytes) {
    memset(hdr, 0, WIRE_HEADER);
    memcpy(hdr, WIRE_MAGIC, 8);
    const uint32_t version = WIRE_VERSION, endian = WIRE_ENDIAN, n = (uint32_t)w->table.size(), start = (uint32_t)WIRE_DATA_START;
    const uint64_t fb = w->cursor;
    const uint32_t tcrc = crc32_update(0, table_bytes.data(), table_bytes.size());
    memcpy(hdr + 8, &version, 4);
    memcpy(hdr + 12, &endian, 4);
    memcpy(hdr + 16, &n, 4);
    memcpy(hdr + 20, &start, 4);
    memcpy(hdr + 24, &fb, 8);
    memcpy(hdr + 32, &tcrc, 4);
}

extern "C" const char* fvm_wire_last_error(fvm_wire_handle w) { return w ? w->err.c_str() : g_wire_err.c_str(); }

extern "C" int32_t fvm_wire_create(const char* path, fvm_wire_handle* out) {
    if (!path || !out) return wire_fail(nullptr, FVM_ERR_ARG, "fvm_wire_create: null argument");
    *out = nullptr;
    FILE* f = fopen(path, "wb");
    if (!f) return wire_fail(nullptr, FVM_ERR_IO, std::string("fvm_wire_create: cannot open ") + path + " for writing");
    fvm_wire* w = new fvm_wire;
    w->f = f;
    w->writing = true;
    w->path = path;
    w->cursor = WIRE_DATA_START;
    // reserve header + table; they are rewritten by fvm_wire_close once every array is known
    std::vector<uint8_t> zeros(WIRE_DATA_START, 0);
    if (fwrite(zeros.data(), 1, zeros.size(), f) != zeros.size()) {
        fclose(f);
        delete w;
        return wire_fail(nullptr, FVM_ERR_IO, "fvm_wire_create: short write");
    }
    *out = w;
    return FVM_OK;
}

extern "C" int32_t fvm_wire_put(fvm_wire_handle w, const char* name, int32_t dtype, int32_t rank, const int64_t* dims,
                                const void* data) {
    if (!w) return wire_fail(nullptr, FVM_ERR_ARG, "fvm_wire_put: null handle");
    if (!w->writing) return wire_fail(w, FVM_ERR_STATE, "fvm_wire_put: container is open for reading");
    if (!name || !name[0] || strlen(name) > 31) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: name must have 1..31 bytes");
    const size_t es = dtype_size(dtype);
    if (!es) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: unknown dtype");
    if (rank < 0 || rank > 4 || (rank > 0 && !dims)) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: rank must be 0..4");
    if (w->table.size() >= FVM_WIRE_MAX_ARRAYS) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: table is full");
    for (const WireEntry& e : w->table)
        if (!strncmp(e.name, name, 32)) return wire_fail(w, FVM_ERR_ARG, std::string("fvm_wire_put: duplicate array name ") + name);
    WireEntry e;
    memset(&e, 0, sizeof e);
    strncpy(e.name, name, 31);
    e.dtype = (uint32_t)dtype;
    e.rank = (uint32_t)rank;
    uint64_t count = 1;
    for (int d = 0; d < rank; ++d) {
        if (dims[d] < 0) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: negative dimension");
        e.dims[d] = dims[d];
        count *= (uint64_t)dims[d];
    }
    e.nbytes = count * es;
    if (e.nbytes && !data) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_put: null data");
    e.offset = w->cursor;
    e.crc = crc32_update(0, data, e.nbytes);
    if (e.nbytes && fwrite(data, 1, e.nbytes, w->f) != e.nbytes) return wire_fail(w, FVM_ERR_IO, "fvm_wire_put: short write");
    const uint64_t end = e.offset + e.nbytes, padded = (end + WIRE_ALIGN - 1) / WIRE_ALIGN * WIRE_ALIGN;
    if (padded > end) {
        const uint8_t zeros[WIRE_ALIGN] = {0};
        if (fwrite(zeros, 1, padded - end, w->f) != padded - end) return wire_fail(w, FVM_ERR_IO, "fvm_wire_put: short write");
    }
    w->cursor = padded;
    w->table.push_back(e);
    return FVM_OK;
}

extern "C" int32_t fvm_wire_open(const char* path, fvm_wire_handle* out) {
    if (!path || !out) return wire_fail(nullptr, FVM_ERR_ARG, "fvm_wire_open: null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return wire_fail(nullptr, FVM_ERR_IO, std::string("fvm_wire_open: cannot open ") + path);
    uint8_t hdr[WIRE_HEADER];
    std::vector<uint8_t> tb(WIRE_ENTRY * FVM_WIRE_MAX_ARRAYS);
    auto bad = [&](const std::string& why) {
        fclose(f);
        return wire_fail(nullptr, FVM_ERR_IO, "fvm_wire_open: " + why + " (" + path + ")");
    };
    if (fread(hdr, 1, WIRE_HEADER, f) != WIRE_HEADER) return bad("file shorter than the header");
    if (memcmp(hdr, WIRE_MAGIC, 8)) return bad("not an FVMWIRE container");
    uint32_t version, endian, n, start, tcrc;
    uint64_t fb;
    memcpy(&version, hdr + 8, 4);
    memcpy(&endian, hdr + 12, 4);
    memcpy(&n, hdr + 16, 4);
    memcpy(&start, hdr + 20, 4);
    memcpy(&fb, hdr + 24, 8);
    memcpy(&tcrc, hdr + 32, 4);
    if (endian != WIRE_ENDIAN) return bad("written with a different byte order");
    if (version != WIRE_VERSION) return bad("unsupported version " + std::to_string(version));
    if (n > FVM_WIRE_MAX_ARRAYS || start != WIRE_DATA_START) return bad("corrupt header");
    if (fread(tb.data(), 1, tb.size(), f) != tb.size()) return bad("file shorter than the array table");
    if (crc32_update(0, tb.data(), tb.size()) != tcrc) return bad("array table checksum mismatch");
    if (fseek(f, 0, SEEK_END)) return bad("seek failed");
    const long sz = ftell(f);
    if (sz < 0 || (uint64_t)sz != fb) return bad("truncated: header says " + std::to_string(fb) + " bytes, file has " + std::to_string(sz));
    fvm_wire* w = new fvm_wire;
    w->f = f;
    w->path = path;
    w->file_bytes = fb;
    w->table.resize(n);
    if (n) memcpy(w->table.data(), tb.data(), n * WIRE_ENTRY);
    for (const WireEntry& e : w->table) {
        uint64_t count = 1;
        for (uint32_t d = 0; d < e.rank && d < 4; ++d) count *= (uint64_t)e.dims[d];
        if (e.rank > 4 || !dtype_size((int32_t)e.dtype) || e.name[31] || count * dtype_size((int32_t)e.dtype) != e.nbytes ||
            e.offset % WIRE_ALIGN || e.offset < WIRE_DATA_START || e.offset + e.nbytes > fb) {
            delete w;
            return bad("corrupt array table entry");
        }
    }
    *out = w;
    return FVM_OK;
}

extern "C" int32_t fvm_wire_count(fvm_wire_handle w, int32_t* n) {
    if (!w || !n) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_count: null argument");
    *n = (int32_t)w->table.size();
    return FVM_OK;
}

extern "C" int32_t fvm_wire_info(fvm_wire_handle w, int32_t index, char* name32, int32_t* dtype, int32_t* rank, int64_t* dims4,
                                 int64_t* nbytes) {
    if (!w) return wire_fail(nullptr, FVM_ERR_ARG, "fvm_wire_info: null handle");
    if (index < 0 || index >= (int32_t)w->table.size()) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_info: index out of range");
    const WireEntry& e = w->table[index];
    if (name32) memcpy(name32, e.name, 32);
    if (dtype) *dtype = (int32_t)e.dtype;
    if (rank) *rank = (int32_t)e.rank;
    if (dims4) memcpy(dims4, e.dims, sizeof e.dims);
    if (nbytes) *nbytes = (int64_t)e.nbytes;
    return FVM_OK;
}

extern "C" int32_t fvm_wire_find(fvm_wire_handle w, const char* name, int32_t* index) {
    if (!w || !name || !index) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_find: null argument");
    *index = -1;
    for (size_t i = 0; i < w->table.size(); ++i)
        if (!strncmp(w->table[i].name, name, 32)) {
            *index = (int32_t)i;
            return FVM_OK;
        }
    return wire_fail(w, FVM_ERR_ARG, std::string("fvm_wire_find: no array named ") + name);
}

extern "C" int32_t fvm_wire_get(fvm_wire_handle w, int32_t index, void* out, int64_t nbytes) {
    if (!w) return wire_fail(nullptr, FVM_ERR_ARG, "fvm_wire_get: null handle");
    if (w->writing) return wire_fail(w, FVM_ERR_STATE, "fvm_wire_get: container is open for writing");
    if (index < 0 || index >= (int32_t)w->table.size()) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_get: index out of range");
    const WireEntry& e = w->table[index];
    if ((uint64_t)nbytes != e.nbytes) return wire_fail(w, FVM_ERR_ARG, std::string("fvm_wire_get: buffer size does not match array ") + e.name);
    if (e.nbytes == 0) return FVM_OK;
    if (!out) return wire_fail(w, FVM_ERR_ARG, "fvm_wire_get: null buffer");
    if (fseek(w->f, (long)e.offset, SEEK_SET)) return wire_fail(w, FVM_ERR_IO, "fvm_wire_get: seek failed");
    if (fread(out, 1, e.nbytes, w->f) != e.nbytes) return wire_fail(w, FVM_ERR_IO, "fvm_wire_get: short read");
    if (crc32_update(0, out, e.nbytes) != e.crc)
        return wire_fail(w, FVM_ERR_IO, std::string("fvm_wire_get: checksum mismatch in array ") + e.name);
    return FVM_OK;
}

extern "C" int32_t fvm_wire_close(fvm_wire_handle w) {
    if (!w) return FVM_OK;
    int32_t rc = FVM_OK;
    if (w->writing) {
        std::vector<uint8_t> tb(WIRE_ENTRY * FVM_WIRE_MAX_ARRAYS, 0);
        if (!w->table.empty()) memcpy(tb.data(), w->table.data(), w->table.size() * WIRE_ENTRY);
        uint8_t hdr[WIRE_HEADER];
        pack_header(w, hdr, tb);
        if (fseek(w->f, 0, SEEK_SET) || fwrite(hdr, 1, WIRE_HEADER, w->f) != WIRE_HEADER || fwrite(tb.data(), 1, tb.size(), w->f) != tb.size())
            rc = wire_fail(nullptr, FVM_ERR_IO, "fvm_wire_close: cannot write the header of " + w->path);
    }
    if (fclose(w->f) && rc == FVM_OK) rc = wire_fail(nullptr, FVM_ERR_IO, "fvm_wire_close: close failed for " + w->path);
    delete w;
    return rc;
}

extern "C" uint32_t fvm_wire_crc32(const void* data, int64_t nbytes) { return crc32_update(0, data, nbytes > 0 ? (size_t)nbytes : 0); }

// ---- engine hook: FVMGeometry(tri) straight from a container ---------------------------------
// arrays read: "points" f64 (2,N), "triangles" i32 (3,T), "index_base" i32 (1) [default 1, Julia],
// optional "boundary_edges" i32 (2,Eb) = keys(get_boundary_edge_map(tri)) in the same index base
extern "C" int32_t fvm_create_from_wire(const char* path, int32_t neq, int32_t device, fvm_handle* out) {
    if (!out) return FVM_ERR_ARG;
    *out = nullptr;
    fvm_wire* w = nullptr;
    int32_t rc = fvm_wire_open(path, &w);
    if (rc) return rc;
    auto fetch = [&](const char* name, int32_t dtype, int64_t lead, std::vector<uint8_t>& buf, int64_t* count, bool required) -> int32_t {
        int32_t idx = -1;
        *count = 0;
        if (fvm_wire_find(w, name, &idx)) return required ? FVM_ERR_ARG : FVM_OK;
        const WireEntry& e = w->table[idx];
        const int64_t have_lead = e.rank >= 2 ? e.dims[0] : 1;
        if ((int32_t)e.dtype != dtype || (lead > 1 && (e.rank != 2 || have_lead != lead)))
            return wire_fail(w, FVM_ERR_ARG, std::string("fvm_create_from_wire: array ") + name + " has the wrong type or shape");
        buf.resize(e.nbytes);
        *count = (int64_t)(e.nbytes / dtype_size(dtype) / (lead > 1 ? lead : 1));
        return fvm_wire_get(w, idx, buf.data(), (int64_t)e.nbytes);
    };
    std::vector<uint8_t> pts, tris, base, bedges;
    int64_t N = 0, T = 0, nb = 0, Eb = 0;
    if ((rc = fetch("points", FVM_WIRE_F64, 2, pts, &N, true)) || (rc = fetch("triangles", FVM_WIRE_I32, 3, tris, &T, true)) ||
        (rc = fetch("index_base", FVM_WIRE_I32, 1, base, &nb, false)) || (rc = fetch("boundary_edges", FVM_WIRE_I32, 2, bedges, &Eb, false))) {
        g_wire_err = w->err;
        fvm_wire_close(w);
        return rc;
    }
    int32_t index_base = 1;
    if (nb >= 1) memcpy(&index_base, base.data(), 4);
    fvm_wire_close(w);
    rc = fvm_create(reinterpret_cast<const double*>(pts.data()), N, reinterpret_cast<const int32_t*>(tris.data()), T, index_base, neq,
                    device, out);
    if (rc) return rc;
    if (Eb > 0) rc = fvm_set_boundary_edges(*out, reinterpret_cast<const int32_t*>(bedges.data()), Eb);
    return rc;
}
