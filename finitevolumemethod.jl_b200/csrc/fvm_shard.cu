// libfvmcuda: multi-GPU sharding.  One process per GPU; every rank owns a node partition plus a
// one-layer ghost halo and ALL triangles touching an owned node (cut triangles are computed
// redundantly on both sides, so `du` never needs exchanging).  Before every operator application
// the packed owned-boundary values of the input vector are exchanged with grouped
// ncclSend/ncclRecv over NVLink and scattered into the ghost slots.
#include <nccl.h>

#include <cstring>

#include "fvm_internal.h"

struct ShardState {
    ncclComm_t comm = nullptr;
    int32_t n_neigh = 0;
    std::vector<int32_t> neigh, send_ptr, recv_ptr;
    int32_t* d_send_idx = nullptr;  // native node ids, grouped by neighbour
    int32_t* d_recv_idx = nullptr;
    double* d_send_buf = nullptr;
    double* d_recv_buf = nullptr;
    uint8_t* d_owned = nullptr;     // native order: 1 = owned (counts in global reductions)
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
    int64_t n_send = 0, n_recv = 0;
};

#define FVM_NCCL(h, call)                                                                            \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return fvm_fail((h), FVM_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

void fvm_shard_release(fvm_ctx* h) {
    ShardState* s = (ShardState*)h->shard;
    if (!s) return;
    if (s->comm) ncclCommDestroy(s->comm);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->ev_packed) cudaEventDestroy(s->ev_packed);
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    delete s;
    h->shard = nullptr;
}

extern "C" int32_t fvm_nccl_unique_id(void* out128) {
    if (!out128) return FVM_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return FVM_ERR_NCCL;
    std::memcpy(out128, &id, sizeof(id));
    return FVM_OK;
}

extern "C" int32_t fvm_shard_init(fvm_handle h, const void* nccl_unique_id, int32_t rank, int32_t nranks) {
    if (!h) return FVM_ERR_ARG;
    if (!h->finalized) return fvm_fail(h, FVM_ERR_STATE, "fvm_shard_init: finalize first");
    FVM_REQUIRE(h, nccl_unique_id && nranks >= 1 && rank >= 0 && rank < nranks, "fvm_shard_init: bad arguments");
    FVM_CUDA(h, cudaSetDevice(h->device));
    if (!h->shard) h->shard = new ShardState();
    ShardState* s = (ShardState*)h->shard;
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    FVM_NCCL(h, ncclCommInitRank(&s->comm, nranks, id, rank));
    h->rank = rank;
    h->nranks = nranks;
    return FVM_OK;
}

// neighbours and the node lists to send / receive, in the caller's (local) node numbering.  Both
// sides must list the shared nodes in the same order (ascending global id).
extern "C" int32_t fvm_set_halo(fvm_handle h, int32_t n_neigh, const int32_t* neigh_ranks, const int32_t* send_ptr,
                                const int32_t* send_nodes, const int32_t* recv_ptr, const int32_t* recv_nodes) {
    if (!h) return FVM_ERR_ARG;
    if (!h->finalized) return fvm_fail(h, FVM_ERR_STATE, "fvm_set_halo: finalize first");
    FVM_REQUIRE(h, n_neigh >= 0 && (n_neigh == 0 || (neigh_ranks && send_ptr && recv_ptr)), "fvm_set_halo: bad arguments");
    FVM_CUDA(h, cudaSetDevice(h->device));
    if (!h->shard) h->shard = new ShardState();
    ShardState* s = (ShardState*)h->shard;
    s->n_neigh = n_neigh;
    s->neigh.assign(neigh_ranks, neigh_ranks + n_neigh);
    s->send_ptr.assign(send_ptr, send_ptr + n_neigh + 1);
    s->recv_ptr.assign(recv_ptr, recv_ptr + n_neigh + 1);
    s->n_send = n_neigh ? send_ptr[n_neigh] : 0;
    s->n_recv = n_neigh ? recv_ptr[n_neigh] : 0;
    std::vector<int32_t> si(s->n_send), ri(s->n_recv);
    for (int64_t k = 0; k < s->n_send; ++k) {
        const int32_t v = send_nodes[k] - h->h_index_base;
        FVM_REQUIRE(h, v >= 0 && v < h->N, "fvm_set_halo: send node out of range");
        si[k] = h->node_new_of_old[v];
    }
    for (int64_t k = 0; k < s->n_recv; ++k) {
        const int32_t v = recv_nodes[k] - h->h_index_base;
        FVM_REQUIRE(h, v >= 0 && v < h->N, "fvm_set_halo: recv node out of range");
        FVM_REQUIRE(h, h->h_ghost.empty() || h->h_ghost[v], "fvm_set_halo: a received node is not flagged as ghost");
        ri[k] = h->node_new_of_old[v];
    }
    int32_t rc;
    if ((rc = fvm_dev_upload(h, &s->d_send_idx, si))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_recv_idx, ri))) return rc;
    if ((rc = fvm_dev_alloc(h, &s->d_send_buf, (size_t)s->n_send * h->neq))) return rc;
    if ((rc = fvm_dev_alloc(h, &s->d_recv_buf, (size_t)s->n_recv * h->neq))) return rc;
    if (!s->comm_stream) {
        FVM_CUDA(h, cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
        FVM_CUDA(h, cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
        FVM_CUDA(h, cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
    }
    h->comm_stream = s->comm_stream;
    // tiles whose local nodes (own range or external interface nodes) contain no received ghost node
    // can run while the exchange is in flight
    {
        std::vector<uint8_t> is_recv(h->N, 0);
        for (int32_t g : ri) is_recv[g] = 1;
        std::vector<int32_t> pre(h->N + 1, 0);
        for (int64_t g = 0; g < h->N; ++g) pre[g + 1] = pre[g] + is_recv[g];
        const int64_t n_tiles = h->dm.n_tiles;
        std::vector<int32_t> indep, dep;
        for (int64_t b = 0; b < n_tiles; ++b) {
            const int32_t n0 = h->h_tile_node0[b], n1 = n0 + h->h_tile_nown[b];
            bool d = pre[n1] - pre[n0] > 0;
            for (int32_t k = h->h_tile_ext0[b]; !d && k < h->h_tile_ext0[b + 1]; ++k) d = is_recv[h->h_ext_ids[k]];
            (d ? dep : indep).push_back((int32_t)b);
        }
        h->n_tiles_indep = (int32_t)indep.size();
        indep.insert(indep.end(), dep.begin(), dep.end());
        if ((rc = fvm_dev_upload(h, &h->d_tile_order, indep))) return rc;
        // measured on 2 B200s at 16.7M nodes per GPU (general kernel): 1 GPU 1.072 ms/step, serialised exchange
        // 1.079 ms, overlapped schedule 1.086 ms.  The 32 KB exchange costs ~7 us over NVLink, so there is
        // nothing left to hide and the extra launches/event waits of the overlapped schedule do not pay:
        // it is opt-in (it matters for small subdomains or many neighbours)
        const char* ov = getenv("FVM_HALO_OVERLAP");
        h->overlap = n_neigh > 0 && ov && ov[0] == '1';
    }
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->halo_ready = true;
    fvm_pipe_release(h);  // a host-buffer pipeline plan made before the halo existed is stale
    return FVM_OK;
}

__global__ void halo_pack_kernel(const int32_t* __restrict__ idx, const int64_t n, const int neq, const double* __restrict__ u,
                                 double* __restrict__ buf) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * neq) return;
    const int64_t q = k / neq;
    buf[k] = u[(int64_t)idx[q] * neq + (k - q * neq)];
}
__global__ void halo_unpack_kernel(const int32_t* __restrict__ idx, const int64_t n, const int neq, const double* __restrict__ buf,
                                   double* __restrict__ u) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * neq) return;
    const int64_t q = k / neq;
    u[(int64_t)idx[q] * neq + (k - q * neq)] = buf[k];
}

// Refreshes the ghost entries of a native-order vector.  Runs on the handle's stream.
int32_t fvm_halo_exchange(fvm_ctx* h, double* u_native) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !h->halo_ready || s->n_neigh == 0) return FVM_OK;
    if (!s->comm) return fvm_fail(h, FVM_ERR_STATE, "halo exchange needs fvm_shard_init");
    const int neq = h->neq;
    if (s->n_send) {
        halo_pack_kernel<<<(unsigned)((s->n_send * neq + 255) / 256), 256, 0, h->stream>>>(s->d_send_idx, s->n_send, neq, u_native,
                                                                                           s->d_send_buf);
        FVM_CUDA(h, cudaGetLastError());
    }
    FVM_NCCL(h, ncclGroupStart());
    for (int q = 0; q < s->n_neigh; ++q) {
        const int64_t ns = (int64_t)(s->send_ptr[q + 1] - s->send_ptr[q]) * neq;
        const int64_t nr = (int64_t)(s->recv_ptr[q + 1] - s->recv_ptr[q]) * neq;
        if (ns) FVM_NCCL(h, ncclSend(s->d_send_buf + (int64_t)s->send_ptr[q] * neq, ns, ncclDouble, s->neigh[q], s->comm, h->stream));
        if (nr) FVM_NCCL(h, ncclRecv(s->d_recv_buf + (int64_t)s->recv_ptr[q] * neq, nr, ncclDouble, s->neigh[q], s->comm, h->stream));
    }
    FVM_NCCL(h, ncclGroupEnd());
    if (s->n_recv) {
        halo_unpack_kernel<<<(unsigned)((s->n_recv * neq + 255) / 256), 256, 0, h->stream>>>(s->d_recv_idx, s->n_recv, neq,
                                                                                             s->d_recv_buf, u_native);
        FVM_CUDA(h, cudaGetLastError());
    }
    return FVM_OK;
}

// Overlapped variant: pack on the compute stream, exchange + unpack on the communication stream.
// The compute stream keeps running kernels that read no ghost entry until fvm_halo_wait().
int32_t fvm_halo_begin(fvm_ctx* h, double* u_native) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !h->halo_ready || s->n_neigh == 0) return FVM_OK;
    if (!s->comm) return fvm_fail(h, FVM_ERR_STATE, "halo exchange needs fvm_shard_init");
    const int neq = h->neq;
    if (s->n_send) {
        halo_pack_kernel<<<(unsigned)((s->n_send * neq + 255) / 256), 256, 0, h->stream>>>(s->d_send_idx, s->n_send, neq, u_native,
                                                                                           s->d_send_buf);
        FVM_CUDA(h, cudaGetLastError());
    }
    FVM_CUDA(h, cudaEventRecord(s->ev_packed, h->stream));
    FVM_CUDA(h, cudaStreamWaitEvent(s->comm_stream, s->ev_packed, 0));
    FVM_NCCL(h, ncclGroupStart());
    for (int q = 0; q < s->n_neigh; ++q) {
        const int64_t ns = (int64_t)(s->send_ptr[q + 1] - s->send_ptr[q]) * neq;
        const int64_t nr = (int64_t)(s->recv_ptr[q + 1] - s->recv_ptr[q]) * neq;
        if (ns) FVM_NCCL(h, ncclSend(s->d_send_buf + (int64_t)s->send_ptr[q] * neq, ns, ncclDouble, s->neigh[q], s->comm, s->comm_stream));
        if (nr) FVM_NCCL(h, ncclRecv(s->d_recv_buf + (int64_t)s->recv_ptr[q] * neq, nr, ncclDouble, s->neigh[q], s->comm, s->comm_stream));
    }
    FVM_NCCL(h, ncclGroupEnd());
    if (s->n_recv) {
        halo_unpack_kernel<<<(unsigned)((s->n_recv * neq + 255) / 256), 256, 0, s->comm_stream>>>(s->d_recv_idx, s->n_recv, neq,
                                                                                                  s->d_recv_buf, u_native);
        FVM_CUDA(h, cudaGetLastError());
    }
    return FVM_OK;
}

int32_t fvm_halo_done(fvm_ctx* h) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !h->halo_ready || s->n_neigh == 0) return FVM_OK;
    FVM_CUDA(h, cudaEventRecord(s->ev_done, s->comm_stream));
    return FVM_OK;
}

int32_t fvm_halo_wait(fvm_ctx* h) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !h->halo_ready || s->n_neigh == 0) return FVM_OK;
    FVM_CUDA(h, cudaStreamWaitEvent(h->stream, s->ev_done, 0));
    return FVM_OK;
}

extern "C" int32_t fvm_halo_exchange_native(fvm_handle h, double* u_native) {
    if (!h || !u_native) return FVM_ERR_ARG;
    if (!h->finalized) return fvm_fail(h, FVM_ERR_STATE, "finalize first");
    FVM_CUDA(h, cudaSetDevice(h->device));
    return fvm_halo_exchange(h, u_native);
}

// sum over ranks of `n` doubles that live on the device (Krylov scalars), on the handle's stream
int32_t fvm_allreduce_sum(fvm_ctx* h, double* d_vals, int n) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !s->comm || h->nranks == 1) return FVM_OK;
    FVM_NCCL(h, ncclAllReduce(d_vals, d_vals, n, ncclDouble, ncclSum, s->comm, h->stream));
    return FVM_OK;
}

// logical OR over ranks of a host flag (e.g. "some rank has Dirichlet nodes"): every rank must take
// the same control-flow decisions, or the halo exchanges no longer pair up
int32_t fvm_global_or(fvm_ctx* h, bool local, bool* global) {
    *global = local;
    ShardState* s = (ShardState*)h->shard;
    if (!s || !s->comm || h->nranks == 1) return FVM_OK;
    double* d = nullptr;
    FVM_CUDA(h, cudaMalloc((void**)&d, sizeof(double)));
    const double v = local ? 1.0 : 0.0;
    double out = 0.0;
    cudaError_t e = cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = ncclAllReduce(d, d, 1, ncclDouble, ncclSum, s->comm, h->stream);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(&out, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (r != ncclSuccess) return fvm_fail(h, FVM_ERR_NCCL, ncclGetErrorString(r));
    FVM_CUDA(h, e);
    *global = out > 0.5;
    return FVM_OK;
}
