// Multi-GPU plumbing: one process per GPU (rank), node-owner partitions, halo refresh of vectors and Krylov
// all-reduces over NCCL/NVLink.  The reference has no distributed path at all; the partition
// analogue is ExtendableGrids' PartitionNodes/PartitionEdges used for threads (src/vfvm_system.jl:741-748).
//
// NCCL is resolved with dlopen at vfvm_comm_init time so that the library loads (and single-GPU runs work) without it.
// The host shares the ncclUniqueId through its own rendezvous (torch.distributed in bench.py / the tests).
#include <dlfcn.h>

#include <cstring>

#include "vfvm_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
    if (g_nccl.lib) return true;
    const char* env = getenv("VFVM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        err = std::string("cannot dlopen libnccl.so.2 (set VFVM_NCCL_LIB): ") + dlerror();
        return false;
    }
#define SYM(field, name)                                        \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);         \
    if (!g_nccl.field) {                                        \
        err = std::string("missing NCCL symbol ") + name;       \
        return false;                                           \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}

__global__ void k_pack(int64_t cnt, int ns, const int32_t* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt * ns) return;
    const int64_t q = i / ns;
    const int s = (int)(i - q * ns);
    buf[i] = x[(int64_t)idx[q] * ns + s];
}

}  // namespace

extern "C" int vfvm_comm_unique_id(char id_out[128]) {
    std::string err;
    if (!load_nccl(err)) return VFVM_ERR_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return VFVM_ERR_COMM;
    memcpy(id_out, id.internal, 128);
    return VFVM_OK;
}

extern "C" int vfvm_comm_init(vfvm_handle* h, int rank, int nranks, const char id[128]) {
    if (!h || rank < 0 || rank >= nranks) return vfvm_fail(h, VFVM_ERR_ARG, "bad rank");
    if (nranks == 1) {
        h->rank = 0;
        h->nranks = 1;
        return VFVM_OK;
    }
    if (!load_nccl(h->err)) return VFVM_ERR_COMM;
    cudaSetDevice(h->device);
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclComm_t comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, nranks, uid, rank);
    if (rc != ncclSuccess) return vfvm_fail(h, VFVM_ERR_COMM, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    h->nccl = comm;
    h->rank = rank;
    h->nranks = nranks;
    return VFVM_OK;
}

int vfvm_comm_destroy(vfvm_handle* h) {
    if (h->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->nccl);
    h->nccl = nullptr;
    return VFVM_OK;
}

extern "C" int vfvm_set_halo(vfvm_handle* h, int nneighbors, const int32_t* neighbor_ranks, const int64_t* send_ptr, const int32_t* send_idx, const int64_t* recv_ptr) {
    if (!h || nneighbors < 0) return vfvm_fail(h, VFVM_ERR_ARG, "bad halo description");
    VFVM_TRY(h, {
        h->nb_ranks.assign(neighbor_ranks, neighbor_ranks + nneighbors);
        h->send_ptr.assign(send_ptr, send_ptr + nneighbors + 1);
        h->recv_ptr.assign(recv_ptr, recv_ptr + nneighbors + 1);
        if (h->recv_ptr[nneighbors] != h->N - h->Nown) return vfvm_fail(h, VFVM_ERR_ARG, "halo ranges do not cover the halo nodes");
        h->send_idx.upload(send_idx, (size_t)h->send_ptr[nneighbors], h->stream);
        h->send_buf.alloc((size_t)std::max<int64_t>(1, h->send_ptr[nneighbors]) * std::max(1, h->n));
        CK(cudaStreamSynchronize(h->stream));
    })
    return VFVM_OK;
}

// refresh x[Nown*n .. N*n) from the owners; x is an n x N device vector
int vfvm_halo_exchange_ptr(vfvm_handle* h, double* x) {
    if (h->nranks <= 1 || h->nb_ranks.empty()) return VFVM_OK;
    const int ns = h->n;
    const int nn = (int)h->nb_ranks.size();
    const int64_t nsend = h->send_ptr[nn];
    if (h->send_buf.n < (size_t)nsend * ns) h->send_buf.alloc((size_t)nsend * ns);
    if (nsend) {
        k_pack<<<cdiv(nsend * ns, 256), 256, 0, h->stream>>>(nsend, ns, h->send_idx.p, x, h->send_buf.p);
        h->launches++;
    }
    ncclComm_t comm = (ncclComm_t)h->nccl;
    g_nccl.GroupStart();
    for (int r = 0; r < nn; r++) {
        const int64_t s0 = h->send_ptr[r], s1 = h->send_ptr[r + 1], r0 = h->recv_ptr[r], r1 = h->recv_ptr[r + 1];
        if (s1 > s0) g_nccl.Send(h->send_buf.p + s0 * ns, (size_t)(s1 - s0) * ns, ncclFloat64, h->nb_ranks[r], comm, h->stream);
        if (r1 > r0) g_nccl.Recv(x + (h->Nown + r0) * ns, (size_t)(r1 - r0) * ns, ncclFloat64, h->nb_ranks[r], comm, h->stream);
    }
    int rc = g_nccl.GroupEnd();
    if (rc != ncclSuccess) throw std::string("NCCL halo exchange: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}

extern "C" int vfvm_halo_exchange(vfvm_handle* h, int which) {
    if (!h || !h->have_pattern || which < 0 || which > 3) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id or no pattern");
    VFVM_TRY(h, {
        vfvm_halo_exchange_ptr(h, h->vec[which].p);
        CK(cudaStreamSynchronize(h->stream));
    })
    return VFVM_OK;
}

int vfvm_comm_allreduce_sum(vfvm_handle* h, double* dev, int count) {
    if (h->nranks <= 1) return VFVM_OK;
    int rc = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}
int vfvm_comm_allreduce_max(vfvm_handle* h, double* dev, int count) {
    if (h->nranks <= 1) return VFVM_OK;
    int rc = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}
