// Multi-GPU: exchange of interface contributions over NCCL (one process per GPU).
// NCCL is bound at run time with dlopen so that single-GPU users do not need it; the handful of entry points used
// here (ncclGetUniqueId, ncclCommInitRank, ncclGroupStart/End, ncclSend, ncclRecv, ncclCommDestroy) have had a
// stable ABI since NCCL 2.7.  The exchange itself is: gather kernel (pack the ghost slots) -> one NCCL group of
// send/recv per peer -> add kernel in ascending peer order (fixed order => reproducible).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "femb_internal.h"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_FLOAT64 = 8 };  // ncclDataType_t ncclFloat64 / ncclDouble

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.handle) return true;
    // FEMB_NCCL_LIB, when set, is the only candidate (a wrong path must fail, not silently bind another NCCL)
    const char* env = getenv("FEMB_NCCL_LIB");
    const char* names[] = {env, env ? nullptr : "libnccl.so.2", env ? nullptr : "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        const char* e = dlerror();  // one call: dlerror() clears the pending message
        g_nccl.err = std::string("cannot dlopen libnccl.so.2 (set FEMB_NCCL_LIB): ") + (e ? e : "");
        return false;
    }
#define BIND(field, sym)                                                      \
    do {                                                                      \
        *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                 \
        if (!g_nccl.field) {                                                  \
            g_nccl.err = std::string("libnccl lacks ") + sym;                 \
            dlclose(g_nccl.handle);                                           \
            g_nccl.handle = nullptr;                                          \
            return false;                                                     \
        }                                                                     \
    } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    return true;
}

#define NCCL_TRY(ctx, expr)                                                                                        \
    do {                                                                                                           \
        int _r = (expr);                                                                                           \
        if (_r != 0) return femb_fail((ctx), FEMB_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r));    \
    } while (0)

__global__ void k_pack(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, double* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// dst[idx[i]] += src[i]; the lists of one peer never repeat a slot, peers are applied in separate launches
__global__ void k_unpack_add(double* __restrict__ dst, const int64_t* __restrict__ idx, int64_t n, const double* __restrict__ src,
                             uint8_t* __restrict__ touched) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const double v = src[i];
        dst[idx[i]] += v;
        if (touched && v != 0.0) touched[idx[i]] = 1;
    }
}

}  // namespace

// drops the exchange plan (the slot lists refer to one particular pattern: called whenever nnz / the slots change)
void femb_free_exchange(femb_ctx* ctx) {
    femb_free(&ctx->send_slots);
    femb_free(&ctx->recv_slots);
    femb_free(&ctx->send_b);
    femb_free(&ctx->recv_b);
    femb_free(&ctx->sendbuf);
    femb_free(&ctx->recvbuf);
    ctx->peers.clear();
    ctx->send_ptr.clear();
    ctx->recv_ptr.clear();
    ctx->sendb_ptr.clear();
    ctx->recvb_ptr.clear();
    ctx->nsend = ctx->nrecv = ctx->nsendb = ctx->nrecvb = 0;
    ctx->xdof_send_lo = ctx->xdof_recv_lo = 0;
    ctx->xdof_send_hi = ctx->xdof_recv_hi = -1;
    ctx->xslices.clear();
    ctx->exchange_version++;
}

int femb_exchange_on(femb_ctx* ctx, cudaStream_t stream) {
    const int np = (int)ctx->peers.size();
    // send buffer layout: [matrix entries of all peers | vector entries of all peers]
    if (ctx->nsend) {
        k_pack<<<(unsigned)((ctx->nsend + 255) / 256), 256, 0, stream>>>(ctx->nzval, ctx->send_slots, ctx->nsend, ctx->sendbuf);
        KERNEL_CHECK(ctx);
    }
    if (ctx->nsendb) {
        k_pack<<<(unsigned)((ctx->nsendb + 255) / 256), 256, 0, stream>>>(ctx->bvec, ctx->send_b, ctx->nsendb, ctx->sendbuf + ctx->nsend);
        KERNEL_CHECK(ctx);
    }
    nccl_comm comm = (nccl_comm)ctx->comm;
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int p = 0; p < np; p++) {
        const int peer = ctx->peers[p];
        const int64_t ns = ctx->send_ptr[p + 1] - ctx->send_ptr[p], nr = ctx->recv_ptr[p + 1] - ctx->recv_ptr[p];
        const int64_t nsb = ctx->sendb_ptr[p + 1] - ctx->sendb_ptr[p], nrb = ctx->recvb_ptr[p + 1] - ctx->recvb_ptr[p];
        if (ns) NCCL_TRY(ctx, g_nccl.Send(ctx->sendbuf + ctx->send_ptr[p], (size_t)ns, NCCL_FLOAT64, peer, comm, stream));
        if (nsb) NCCL_TRY(ctx, g_nccl.Send(ctx->sendbuf + ctx->nsend + ctx->sendb_ptr[p], (size_t)nsb, NCCL_FLOAT64, peer, comm, stream));
        if (nr) NCCL_TRY(ctx, g_nccl.Recv(ctx->recvbuf + ctx->recv_ptr[p], (size_t)nr, NCCL_FLOAT64, peer, comm, stream));
        if (nrb) NCCL_TRY(ctx, g_nccl.Recv(ctx->recvbuf + ctx->nrecv + ctx->recvb_ptr[p], (size_t)nrb, NCCL_FLOAT64, peer, comm, stream));
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    for (int p = 0; p < np; p++) {  // ascending peer order: the sum order of every slot is fixed
        const int64_t nr = ctx->recv_ptr[p + 1] - ctx->recv_ptr[p], nrb = ctx->recvb_ptr[p + 1] - ctx->recvb_ptr[p];
        if (nr) {
            k_unpack_add<<<(unsigned)((nr + 255) / 256), 256, 0, stream>>>(ctx->nzval, ctx->recv_slots + ctx->recv_ptr[p], nr, ctx->recvbuf + ctx->recv_ptr[p],
                                                                         ctx->touched);
            KERNEL_CHECK(ctx);
        }
        if (nrb) {
            k_unpack_add<<<(unsigned)((nrb + 255) / 256), 256, 0, stream>>>(ctx->bvec, ctx->recv_b + ctx->recvb_ptr[p], nrb,
                                                                          ctx->recvbuf + ctx->nrecv + ctx->recvb_ptr[p], nullptr);
            KERNEL_CHECK(ctx);
        }
    }
    return FEMB_OK;
}

extern "C" {

int32_t femb_comm_unique_id(char id_out[128]) {
    if (!id_out) return FEMB_ERR_ARG;
    if (!load_nccl()) return FEMB_ERR_NCCL;
    nccl_uid id;
    if (g_nccl.GetUniqueId(&id) != 0) return FEMB_ERR_NCCL;
    memcpy(id_out, id.internal, 128);
    return FEMB_OK;
}

int32_t femb_comm_init(femb_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]) {
    FEMB_CHECK_CTX(ctx);
    if (nranks < 1 || rank < 0 || rank >= nranks || !id) return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_init: bad arguments");
    if (!load_nccl()) return femb_fail(ctx, FEMB_ERR_NCCL, "%s", g_nccl.err.c_str());
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->comm) femb_comm_destroy(ctx);
    nccl_uid uid;
    memcpy(uid.internal, id, 128);
    nccl_comm comm = nullptr;
    NCCL_TRY(ctx, g_nccl.CommInitRank(&comm, nranks, uid, rank));
    ctx->comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return FEMB_OK;
}

int32_t femb_comm_set_exchange(femb_ctx* ctx, int32_t npeers, const int32_t* peers, const int64_t* send_ptr, const int64_t* send_slots,
                               const int64_t* recv_ptr, const int64_t* recv_slots, const int64_t* sendb_ptr, const int64_t* send_b,
                               const int64_t* recvb_ptr, const int64_t* recv_b) {
    FEMB_CHECK_CTX(ctx);
    if (npeers < 0 || (npeers && (!peers || !send_ptr || !recv_ptr || !sendb_ptr || !recvb_ptr)))
        return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_set_exchange: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // everything is validated on locals first; the context only changes once every list has been checked and copied
    if (npeers == 0) {
        femb_free_exchange(ctx);
        return FEMB_OK;
    }
    for (int p = 0; p < npeers; p++) {
        if (peers[p] < 0 || peers[p] >= ctx->nranks || peers[p] == ctx->rank) return femb_fail(ctx, FEMB_ERR_ARG, "peer %d is not a valid remote rank", peers[p]);
        if (p && peers[p] <= peers[p - 1]) return femb_fail(ctx, FEMB_ERR_ARG, "peers must be ascending (fixed summation order)");
    }
    const int64_t* ptrs[4] = {send_ptr, recv_ptr, sendb_ptr, recvb_ptr};
    for (const int64_t* q : ptrs) {
        if (q[0] != 0) return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_set_exchange: pointer arrays must start at 0");
        for (int p = 0; p < npeers; p++)
            if (q[p + 1] < q[p]) return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_set_exchange: pointer arrays must be non-decreasing");
    }
    const int64_t nsend = send_ptr[npeers], nrecv = recv_ptr[npeers], nsendb = sendb_ptr[npeers], nrecvb = recvb_ptr[npeers];
    if ((nsend && !send_slots) || (nrecv && !recv_slots) || (nsendb && !send_b) || (nrecvb && !recv_b))
        return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_set_exchange: a slot list is NULL but its count is not zero");
    if ((nsend || nrecv) && !ctx->nzval) return femb_fail(ctx, FEMB_ERR_ARG, "femb_comm_set_exchange: build or adopt the pattern first");
    // range checks on the host (the lists are O(interface))
    for (int64_t i = 0; i < nsend; i++)
        if (send_slots[i] < 0 || send_slots[i] >= ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "send slot out of range");
    for (int64_t i = 0; i < nrecv; i++)
        if (recv_slots[i] < 0 || recv_slots[i] >= ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "recv slot out of range");
    for (int64_t i = 0; i < nsendb; i++)
        if (send_b[i] < 0 || send_b[i] >= ctx->nrows) return femb_fail(ctx, FEMB_ERR_ARG, "send vector index out of range");
    for (int64_t i = 0; i < nrecvb; i++)
        if (recv_b[i] < 0 || recv_b[i] >= ctx->nrows) return femb_fail(ctx, FEMB_ERR_ARG, "recv vector index out of range");
    femb_free_exchange(ctx);
    int st = femb_alloc(ctx, &ctx->send_slots, (size_t)nsend);
    if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->recv_slots, (size_t)nrecv);
    if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->send_b, (size_t)nsendb);
    if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->recv_b, (size_t)nrecvb);
    if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->sendbuf, (size_t)(nsend + nsendb));
    if (st == FEMB_OK) st = femb_alloc(ctx, &ctx->recvbuf, (size_t)(nrecv + nrecvb));
    if (st == FEMB_OK) st = femb_h2d(ctx, ctx->send_slots, send_slots, (size_t)nsend);
    if (st == FEMB_OK) st = femb_h2d(ctx, ctx->recv_slots, recv_slots, (size_t)nrecv);
    if (st == FEMB_OK) st = femb_h2d(ctx, ctx->send_b, send_b, (size_t)nsendb);
    if (st == FEMB_OK) st = femb_h2d(ctx, ctx->recv_b, recv_b, (size_t)nrecvb);
    if (st == FEMB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = femb_fail(ctx, FEMB_ERR_CUDA, "femb_comm_set_exchange: copy failed");
    if (st != FEMB_OK) {
        femb_free_exchange(ctx);  // no half-installed plan
        return st;
    }
    ctx->peers.assign(peers, peers + npeers);
    ctx->send_ptr.assign(send_ptr, send_ptr + npeers + 1);
    ctx->recv_ptr.assign(recv_ptr, recv_ptr + npeers + 1);
    ctx->sendb_ptr.assign(sendb_ptr, sendb_ptr + npeers + 1);
    ctx->recvb_ptr.assign(recvb_ptr, recvb_ptr + npeers + 1);
    ctx->nsend = nsend;
    ctx->nrecv = nrecv;
    ctx->nsendb = nsendb;
    ctx->nrecvb = nrecvb;
    if (nsendb) {
        ctx->xdof_send_lo = *std::min_element(send_b, send_b + nsendb);
        ctx->xdof_send_hi = *std::max_element(send_b, send_b + nsendb);
    }
    if (nrecvb) {
        ctx->xdof_recv_lo = *std::min_element(recv_b, recv_b + nrecvb);
        ctx->xdof_recv_hi = *std::max_element(recv_b, recv_b + nrecvb);
    }
    ctx->xslices.clear();
    for (int64_t i = 0; i < nsendb; i++) ctx->xslices.push_back(send_b[i] >> 5);
    for (int64_t i = 0; i < nrecvb; i++) ctx->xslices.push_back(recv_b[i] >> 5);
    std::sort(ctx->xslices.begin(), ctx->xslices.end());
    ctx->xslices.erase(std::unique(ctx->xslices.begin(), ctx->xslices.end()), ctx->xslices.end());
    ctx->exchange_version++;
    ctx->exchange_nnz = ctx->nnz;
    return FEMB_OK;
}

int32_t femb_exchange(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    if (!ctx->comm) return femb_fail(ctx, FEMB_ERR_ARG, "femb_exchange: femb_comm_init has not been called");
    if (ctx->peers.empty()) return FEMB_OK;
    if (!ctx->nzval || !ctx->bvec) return femb_fail(ctx, FEMB_ERR_ARG, "femb_exchange: no matrix / vector");
    if (ctx->exchange_nnz != ctx->nnz) return femb_fail(ctx, FEMB_ERR_ARG, "femb_exchange: the exchange plan was built for another pattern");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FEMB_TRY(femb_flush_zero(ctx, true, true));
    FEMB_TRY(femb_timer_begin(ctx));
    FEMB_TRY(femb_exchange_on(ctx, ctx->stream));
    return femb_timer_end(ctx);
}

int32_t femb_set_fused_exchange(femb_ctx* ctx, int32_t enable) {
    FEMB_CHECK_CTX(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (enable && !ctx->stream_x) {
        int lo = 0, hi = 0;  // highest priority: the few boundary CTAs should not queue behind the bulk grid
        CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->stream_x, cudaStreamNonBlocking, hi));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_x0, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_x1, cudaEventDisableTiming));
    }
    ctx->fused_exchange = enable != 0;
    return FEMB_OK;
}

int32_t femb_comm_destroy(femb_ctx* ctx) {
    FEMB_CHECK_CTX(ctx);
    femb_free_exchange(ctx);
    if (ctx->stream_x) {
        cudaStreamDestroy(ctx->stream_x);
        cudaEventDestroy(ctx->ev_x0);
        cudaEventDestroy(ctx->ev_x1);
        ctx->stream_x = nullptr;
    }
    ctx->fused_exchange = false;
    if (ctx->comm && g_nccl.handle) g_nccl.CommDestroy((nccl_comm)ctx->comm);
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return FEMB_OK;
}

}  // extern "C"
