// libfvmcuda: multi-GPU sharding.  One process per GPU; every rank owns a node partition plus a
// one-layer ghost halo and ALL triangles touching an owned node (cut triangles are computed
// redundantly on both sides, so `du` never needs exchanging).  Before every operator application
// the packed owned-boundary values of the input vector are exchanged with grouped
// ncclSend/ncclRecv over NVLink and scattered into the ghost slots.
#include <nccl.h>

#include <algorithm>
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
    // ---- peer-mapped exchange (NVLink direct stores instead of pack -> ncclSend/ncclRecv -> unpack) ----------------
    // Every rank owns a receive slab: two parities x n_recv x neq doubles, plus one 64-bit epoch flag per (parity,
    // neighbour).  Neighbours map it through CUDA IPC.  An exchange is ONE kernel that gathers the owned boundary values
    // and stores them straight into the neighbours' slabs (the last CTA to finish publishes the epoch to every neighbour
    // with a system-scope release) and ONE kernel that waits for the neighbours' epochs and scatters the slab into the
    // ghost slots.  Parities alternate, so a fast neighbour can already push exchange e+1 while this rank still unpacks e
    // (it cannot get to e+2 before it has seen this rank's push of e+1, which is issued after this rank's unpack of e).
    bool peer = false;
    double* slab = nullptr;                   // [2][n_recv * neq]
    unsigned long long* flags = nullptr;      // [2][n_neigh], in the same allocation as the slab
    void* slab_base = nullptr;
    std::vector<void*> peer_base;             // mapped slab of neighbour q
    double** d_peer_data = nullptr;           // [n_neigh] where my segment starts in neighbour q's slab (parity 0)
    unsigned long long** d_peer_flag = nullptr;  // [n_neigh] my flag in neighbour q's slab (parity 0)
    int64_t* d_peer_stride = nullptr;         // [n_neigh] doubles between neighbour q's parities
    int32_t* d_peer_nneigh = nullptr;         // [n_neigh] flags per parity in neighbour q's slab
    int32_t* d_send_ptr = nullptr;            // [n_neigh + 1]
    int32_t* d_recv_ptr = nullptr;
    unsigned int* d_done_ctr = nullptr;       // CTA completion counter of the push kernel
    int32_t* d_timeout = nullptr;             // set by the wait kernel if a neighbour never signals
    unsigned long long epoch = 0;
};

#define FVM_NCCL(h, call)                                                                            \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return fvm_fail((h), FVM_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

static int32_t peer_setup(fvm_ctx* h, ShardState* s);

void fvm_shard_release(fvm_ctx* h) {
    ShardState* s = (ShardState*)h->shard;
    if (!s) return;
    for (void* p : s->peer_base)
        if (p) cudaIpcCloseMemHandle(p);
    if (s->slab_base) cudaFree(s->slab_base);
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
    if ((rc = peer_setup(h, s))) return rc;
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
        // Overlapped schedule (exchange + halo-dependent tiles on the communication stream, independent tiles on the
        // compute stream).  With the NCCL exchange it did not pay at 16.7M nodes per GPU (round 1, general kernel: 1 GPU
        // 1.072 ms/step, serialised 1.079 ms, overlapped 1.086 ms).  With the peer-mapped exchange it does (2 B200s,
        // gpurun_out/r2w_overlap_ab.log: RHS 0.373 -> 0.364 ms, SpMV 0.294 -> 0.285 ms, Tsit5 2.43 -> 2.37 ms per step),
        // so it is the default there; FVM_HALO_OVERLAP=0 / 1 overrides.
        const char* ov = getenv("FVM_HALO_OVERLAP");
        h->overlap = n_neigh > 0 && (ov ? ov[0] == '1' : s->peer);
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


// ---- peer-mapped exchange ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    halo_push_kernel(const int32_t* __restrict__ send_idx, const int32_t* __restrict__ send_ptr, const int n_neigh, const int neq,
                     const double* __restrict__ u, double* const* __restrict__ peer_data, unsigned long long* const* __restrict__ peer_flag,
                     const int64_t* __restrict__ peer_stride, const int32_t* __restrict__ peer_nneigh, const int parity,
                     const unsigned long long epoch, unsigned int* __restrict__ done_ctr) {
    const int64_t n = (int64_t)send_ptr[n_neigh] * neq;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = k / neq;
        const int v = (int)(k - e * neq);
        int q = 0;
        while (e >= send_ptr[q + 1]) ++q;  // a handful of neighbours
        peer_data[q][(int64_t)parity * peer_stride[q] + (e - send_ptr[q]) * neq + v] = u[(int64_t)send_idx[e] * neq + v];
    }
    // the last CTA to get here publishes the epoch: every CTA's stores are fenced at system scope before it counts itself
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned int last;
    if (threadIdx.x == 0) last = atomicAdd(done_ctr, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (last) {
        __threadfence_system();
        for (int q = threadIdx.x; q < n_neigh; q += blockDim.x) {
            unsigned long long* f = peer_flag[q] + (int64_t)parity * peer_nneigh[q];
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
        if (threadIdx.x == 0) *done_ctr = 0;
    }
}

// one or more CTAs per neighbour: wait for its epoch, then scatter its slab segment into the ghost slots
__global__ void __launch_bounds__(256)
    halo_wait_unpack_kernel(const int32_t* __restrict__ recv_idx, const int32_t* __restrict__ recv_ptr, const int n_neigh, const int neq,
                            const double* __restrict__ slab, const unsigned long long* __restrict__ flags, const int64_t slab_stride,
                            const int parity, const unsigned long long epoch, double* __restrict__ u, int32_t* __restrict__ timeout_flag) {
    const int q = blockIdx.y;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned long long* f = flags + (int64_t)parity * n_neigh + q;
        unsigned long long seen = 0;
        const long long t0 = clock64();
        ok = 1;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
            if (seen >= epoch) break;
            if (clock64() - t0 > 20000000000LL) {  // ~10 s: a neighbour died; fail loudly instead of hanging the device
                ok = 0;
                atomicExch(timeout_flag, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!ok) return;
    const int64_t lo = (int64_t)recv_ptr[q] * neq, hi = (int64_t)recv_ptr[q + 1] * neq;
    const double* src = slab + (int64_t)parity * slab_stride;
    for (int64_t k = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = k / neq;
        u[(int64_t)recv_idx[e] * neq + (k - e * neq)] = __ldcg(src + k);
    }
}

// maps the neighbours' receive slabs (CUDA IPC); on any failure the NCCL path stays in use
static int32_t peer_setup(fvm_ctx* h, ShardState* s) {
    s->peer = false;
    const char* env = getenv("FVM_HALO_PEER");
    int want = (env && env[0] == '0') ? 0 : 1;
    if (!s->comm || h->nranks < 2) return FVM_OK;
    const int neq = h->neq, nr = h->nranks;
    const int64_t stride = std::max<int64_t>(2, (s->n_recv * neq + 1) & ~(int64_t)1);
    struct Info {
        cudaIpcMemHandle_t handle;
        int64_t stride;
        int32_t n_neigh, ok;
    };
    Info mine{};
    mine.stride = stride;
    mine.n_neigh = s->n_neigh;
    mine.ok = 0;
    void* base = nullptr;
    const size_t bytes = sizeof(double) * 2 * stride + sizeof(unsigned long long) * 2 * std::max(1, s->n_neigh);
    if (want && cudaMalloc(&base, bytes) == cudaSuccess && cudaMemset(base, 0, bytes) == cudaSuccess &&
        cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess)
        mine.ok = 1;
    cudaGetLastError();
    // everyone learns everyone's handle and, per pair, where the sender's segment and flag live in the receiver's slab
    std::vector<Info> all(nr);
    std::vector<int64_t> my_off(nr, -1), all_off((size_t)nr * nr, -1);  // my_off[q]: offset (doubles) of q's segment in MY slab
    std::vector<int32_t> my_slot(nr, -1), all_slot((size_t)nr * nr, -1);
    for (int q = 0; q < s->n_neigh; ++q) {
        my_off[s->neigh[q]] = (int64_t)s->recv_ptr[q] * neq;
        my_slot[s->neigh[q]] = q;
    }
    char* d_tmp = nullptr;
    const size_t per = sizeof(Info) + sizeof(int64_t) * nr + sizeof(int32_t) * nr;
    FVM_CUDA(h, cudaMalloc((void**)&d_tmp, per * (nr + 1)));
    std::vector<char> pack(per), gathered(per * nr);
    std::memcpy(pack.data(), &mine, sizeof(Info));
    std::memcpy(pack.data() + sizeof(Info), my_off.data(), sizeof(int64_t) * nr);
    std::memcpy(pack.data() + sizeof(Info) + sizeof(int64_t) * nr, my_slot.data(), sizeof(int32_t) * nr);
    cudaMemcpyAsync(d_tmp, pack.data(), per, cudaMemcpyHostToDevice, h->stream);
    ncclResult_t nres = ncclAllGather(d_tmp, d_tmp + per, per, ncclChar, s->comm, h->stream);
    cudaMemcpyAsync(gathered.data(), d_tmp + per, per * nr, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    cudaFree(d_tmp);
    if (nres != ncclSuccess) return fvm_fail(h, FVM_ERR_NCCL, ncclGetErrorString(nres));
    FVM_CUDA(h, ce);
    bool all_ok = true;
    for (int r = 0; r < nr; ++r) {
        std::memcpy(&all[r], gathered.data() + per * r, sizeof(Info));
        std::memcpy(all_off.data() + (size_t)r * nr, gathered.data() + per * r + sizeof(Info), sizeof(int64_t) * nr);
        std::memcpy(all_slot.data() + (size_t)r * nr, gathered.data() + per * r + sizeof(Info) + sizeof(int64_t) * nr, sizeof(int32_t) * nr);
        all_ok = all_ok && all[r].ok;
    }
    std::vector<void*> mapped(s->n_neigh, nullptr);
    int my_ok = all_ok ? 1 : 0;
    for (int q = 0; q < s->n_neigh && my_ok; ++q) {
        const int r = s->neigh[q];
        if (all_off[(size_t)r * nr + h->rank] < 0 || cudaIpcOpenMemHandle(&mapped[q], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            my_ok = 0;
        }
    }
    // the choice must be unanimous: a rank on the NCCL path cannot pair with one that pushes
    double* d_flag = nullptr;
    FVM_CUDA(h, cudaMalloc((void**)&d_flag, sizeof(double)));
    const double fv = my_ok ? 0.0 : 1.0;
    double bad = 0.0;
    cudaMemcpyAsync(d_flag, &fv, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    nres = ncclAllReduce(d_flag, d_flag, 1, ncclDouble, ncclSum, s->comm, h->stream);
    cudaMemcpyAsync(&bad, d_flag, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    ce = cudaStreamSynchronize(h->stream);
    cudaFree(d_flag);
    if (nres != ncclSuccess) return fvm_fail(h, FVM_ERR_NCCL, ncclGetErrorString(nres));
    FVM_CUDA(h, ce);
    if (bad > 0.5) {
        for (void* p : mapped)
            if (p) cudaIpcCloseMemHandle(p);
        if (base) cudaFree(base);
        cudaGetLastError();
        return FVM_OK;  // NCCL send/recv stays
    }
    s->slab_base = base;
    s->slab = (double*)base;
    s->flags = (unsigned long long*)((char*)base + sizeof(double) * 2 * stride);
    s->peer_base = mapped;
    std::vector<double*> pd(s->n_neigh);
    std::vector<unsigned long long*> pf(s->n_neigh);
    std::vector<int64_t> ps(s->n_neigh);
    std::vector<int32_t> pn(s->n_neigh);
    for (int q = 0; q < s->n_neigh; ++q) {
        const int r = s->neigh[q];
        pd[q] = (double*)mapped[q] + all_off[(size_t)r * nr + h->rank];
        pf[q] = (unsigned long long*)((char*)mapped[q] + sizeof(double) * 2 * all[r].stride) + all_slot[(size_t)r * nr + h->rank];
        ps[q] = all[r].stride;
        pn[q] = all[r].n_neigh;
    }
    int32_t rc;
    if ((rc = fvm_dev_upload(h, &s->d_peer_data, pd))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_peer_flag, pf))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_peer_stride, ps))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_peer_nneigh, pn))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_send_ptr, s->send_ptr))) return rc;
    if ((rc = fvm_dev_upload(h, &s->d_recv_ptr, s->recv_ptr))) return rc;
    if ((rc = fvm_dev_alloc(h, &s->d_done_ctr, 1))) return rc;
    if ((rc = fvm_dev_alloc(h, &s->d_timeout, 1))) return rc;
    FVM_CUDA(h, cudaMemsetAsync(s->d_done_ctr, 0, sizeof(unsigned int), h->stream));
    FVM_CUDA(h, cudaMemsetAsync(s->d_timeout, 0, sizeof(int32_t), h->stream));
    FVM_CUDA(h, cudaStreamSynchronize(h->stream));
    s->epoch = 0;
    s->peer = true;
    return FVM_OK;
}

// both halves of a peer exchange on the given streams (the same stream for the serialised schedule)
static int32_t peer_exchange(fvm_ctx* h, ShardState* s, double* u_native, cudaStream_t push_stream, cudaStream_t wait_stream) {
    const int neq = h->neq;
    s->epoch += 1;
    const int parity = (int)(s->epoch & 1);
    const int64_t stride = std::max<int64_t>(2, (s->n_recv * neq + 1) & ~(int64_t)1);
    if (s->n_neigh > 0) {
        const int64_t n = s->n_send * neq;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(296, (n + 255) / 256));
        halo_push_kernel<<<grid, 256, 0, push_stream>>>(s->d_send_idx, s->d_send_ptr, s->n_neigh, neq, u_native, s->d_peer_data, s->d_peer_flag,
                                                        s->d_peer_stride, s->d_peer_nneigh, parity, s->epoch, s->d_done_ctr);
        FVM_CUDA(h, cudaGetLastError());
        int64_t longest = 1;
        for (int q = 0; q < s->n_neigh; ++q) longest = std::max<int64_t>(longest, (int64_t)(s->recv_ptr[q + 1] - s->recv_ptr[q]) * neq);
        dim3 g((unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (longest + 255) / 256)), (unsigned)s->n_neigh);
        halo_wait_unpack_kernel<<<g, 256, 0, wait_stream>>>(s->d_recv_idx, s->d_recv_ptr, s->n_neigh, neq, s->slab, s->flags, stride, parity,
                                                            s->epoch, u_native, s->d_timeout);
        FVM_CUDA(h, cudaGetLastError());
    }
    return FVM_OK;
}

// Refreshes the ghost entries of a native-order vector.  Runs on the handle's stream.
int32_t fvm_halo_exchange(fvm_ctx* h, double* u_native) {
    ShardState* s = (ShardState*)h->shard;
    if (!s || !h->halo_ready || s->n_neigh == 0) return FVM_OK;
    if (!s->comm) return fvm_fail(h, FVM_ERR_STATE, "halo exchange needs fvm_shard_init");
    if (s->peer) return peer_exchange(h, s, u_native, h->stream, h->stream);
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
    if (s->peer) {  // push on the compute stream, wait + unpack on the communication stream
        FVM_CUDA(h, cudaEventRecord(s->ev_packed, h->stream));
        FVM_CUDA(h, cudaStreamWaitEvent(s->comm_stream, s->ev_packed, 0));  // u_native is complete
        return peer_exchange(h, s, u_native, s->comm_stream, s->comm_stream);
    }
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

// 0: no halo / single rank, 1: NCCL send/recv, 2: peer-mapped NVLink stores.  *timed_out != 0: a neighbour never signalled
extern "C" int32_t fvm_halo_mode(fvm_handle h, int32_t* mode, int32_t* timed_out) {
    if (!h || !mode) return FVM_ERR_ARG;
    ShardState* s = (ShardState*)h->shard;
    *mode = (!s || !h->halo_ready || s->n_neigh == 0) ? 0 : (s->peer ? 2 : 1);
    if (timed_out) {
        *timed_out = 0;
        if (s && s->peer) {
            FVM_CUDA(h, cudaSetDevice(h->device));
            FVM_CUDA(h, cudaMemcpy(timed_out, s->d_timeout, sizeof(int32_t), cudaMemcpyDeviceToHost));
        }
    }
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
