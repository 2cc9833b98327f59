// Multi-GPU plumbing: one process per GPU (rank), node-owner partitions, halo refresh of vectors and Krylov
// all-reduces over NCCL/NVLink.  The reference has no distributed path at all; the partition
// analogue is ExtendableGrids' PartitionNodes/PartitionEdges used for threads (src/vfvm_system.jl:741-748).
//
// NCCL is resolved with dlopen at vfvm_comm_init time so that the library loads (and single-GPU runs work) without it.
// The host shares the ncclUniqueId through its own rendezvous (torch.distributed in bench.py / the tests).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "vfvm_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat64 = 8 };  // ncclDataType_t: ncclChar, ncclDouble
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllGather, "ncclAllGather")
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

__global__ void k_clear_bits(int32_t* w, int32_t bits) { atomicAnd(w, ~bits); }

// ---- peer mailboxes (peer.cuh) ---------------------------------------------------------------------------------------
struct BoxLayout {
    size_t dir, hflag, rflag, red, halo;
};
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// the header part depends on the number of ranks only, so every rank can address every other rank's mailbox
BoxLayout box_layout(int R) {
    BoxLayout b;
    b.dir = 0;  // int64 recv_off[R], slot[R], halo_doubles
    b.hflag = align256(b.dir + sizeof(int64_t) * (2 * R + 1));             // u64 [2][R]
    b.rflag = b.hflag + sizeof(uint64_t) * 2 * R;                           // u64 [2][R]
    b.red = align256(b.rflag + sizeof(uint64_t) * 2 * R);                   // double [2][R][VFVM_PEER_RED_W]
    b.halo = align256(b.red + sizeof(double) * 2 * R * VFVM_PEER_RED_W);    // double [2][halo_doubles]
    return b;
}

// generic halo refresh of a vector through the mailboxes, one kernel: push, raise flags, wait for the neighbours, unpack
template <int NS>
__global__ void k_peer_halo(const PeerArgs P, double* __restrict__ x) {
    const unsigned long long seq = peer_seq(P);
    peer_push<NS>(P, seq, x);
    if (threadIdx.x < P.nn) peer_wait(peer_hflag_local(P, seq) + threadIdx.x, seq, P.err, P.timeout_ns);
    __syncthreads();
    const int64_t total = P.nhalo * NS;
    const double* __restrict__ box = peer_halo_local(P, seq);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / NS;
        x[P.Nown * NS + i] = peer_ld_data(box + peer_halo_pos(P, c) * NS + (i - c * NS));
    }
}

// in-place all-reduce (sum or max) of `count` <= VFVM_PEER_RED_W device values: store my values into every rank's box, wait
// for every rank's values in mine, combine in rank order
__global__ void k_peer_allreduce(const PeerArgs P, double* __restrict__ vals, int count, int ismax) {
    const int t = threadIdx.x;
    const unsigned long long seq = peer_seq(P);
    const double* red = peer_reduce_exchange(P, seq, vals, count);
    if (t < count) {
        double acc = peer_ld_data(red + t);
        for (int q = 1; q < P.nranks; q++) {
            const double v = peer_ld_data(red + (size_t)q * VFVM_PEER_RED_W + t);
            acc = ismax ? fmax(acc, v) : acc + v;
        }
        vals[t] = acc;
    }
}


// all-gather of vec = [nranks][cap] in place through the gather boxes: push my segment into every peer's box, raise my flag there, wait
// for the peers' flags, copy their segments out of my box.  One kernel; the grid is one resident wave (the block that finishes the push
// last raises the flags).  16-byte stores over NVLink (cap is even).
__global__ void k_peer_allgather(const GatherArgs G, double* __restrict__ vec) {
    const unsigned long long seq = *(volatile unsigned long long*)G.seq_ctr + 1ull;
    const int par = (int)(seq & 1ull);
    const int R = G.nranks;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t cap2 = G.cap >> 1;
    const double2* __restrict__ mine = (const double2*)(vec + (int64_t)G.rank * G.cap);
    for (int q = 0; q < R; q++) {
        if (q == G.rank) continue;
        double2* dst = (double2*)(G.dst[q] + ((int64_t)par * R + G.rank) * G.cap);
        for (int64_t i = tid; i < cap2; i += nth) dst[i] = mine[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(G.count, 1u);
        if (prev == gridDim.x - 1) {
            *G.count = 0;
            *G.seq_ctr = seq;
            __threadfence_system();
            for (int q = 0; q < R; q++)
                if (q != G.rank) peer_st_flag(G.flag_dst[q] + par * R + G.rank, seq);
        }
    }
    if ((int)threadIdx.x < R && (int)threadIdx.x != G.rank) peer_wait(G.flag_local + par * R + threadIdx.x, seq, G.err, G.timeout_ns);
    __syncthreads();
    for (int q = 0; q < R; q++) {
        if (q == G.rank) continue;
        const double* __restrict__ src = G.box_local + ((int64_t)par * R + q) * G.cap;
        double* __restrict__ out = vec + (int64_t)q * G.cap;
        for (int64_t i = tid; i < G.cap; i += nth) out[i] = peer_ld_data(src + i);
    }
}

}  // namespace

static void fill_peer_common(vfvm_handle* h, PeerArgs& P) {
    memset(&P, 0, sizeof(P));
    P.nn = (int)h->nb_ranks.size();
    P.nranks = h->nranks;
    P.rank = h->rank;
    P.ns = h->n;
    for (int r = 0; r <= P.nn; r++) {
        P.send_ptr[r] = h->send_ptr[r];
        P.recv_ptr0[r] = P.recv_ptrl[r] = h->recv_ptr[r];
    }
    P.send_idx = h->send_idx.p;
    P.Nown = h->Nown;
    P.nhalo = h->N - h->Nown;
    P.push_count = h->peer_counter.p;
    P.err = h->flags.p;
    P.timeout_ns = h->peer_timeout_ns;
}

PeerArgs vfvm_peer_args_halo(vfvm_handle* h) {
    PeerArgs P;
    fill_peer_common(h, P);
    P.seq_ctr = h->peer_seq.p;
    const int R = h->nranks;
    const BoxLayout b = box_layout(R);
    for (int r = 0; r < P.nn; r++) {
        char* base = h->peer_base[h->nb_ranks[r]];
        P.halo_dst[r] = (double*)(base + b.halo) + (size_t)h->peer_recv_off[r] * h->n;
        P.halo_dst_stride[r] = h->peer_halo_doubles[r];
        P.hflag_dst[r] = (unsigned long long*)(base + b.hflag) + h->peer_slot[r];
    }
    P.halo_local = (const double*)(h->peer_box + b.halo);
    P.halo_local_stride = (h->N - h->Nown) * h->n;
    P.hflag_local = (const unsigned long long*)(h->peer_box + b.hflag);
    return P;
}

PeerArgs vfvm_peer_args_reduce(vfvm_handle* h) {
    PeerArgs P;
    fill_peer_common(h, P);
    P.seq_ctr = h->peer_seq.p + 1;
    const int R = h->nranks;
    const BoxLayout b = box_layout(R);
    for (int q = 0; q < R; q++) {
        char* base = h->peer_base[q];
        P.red_dst[q] = (double*)(base + b.red) + (size_t)h->rank * VFVM_PEER_RED_W;
        P.rflag_dst[q] = (unsigned long long*)(base + b.rflag) + h->rank;
    }
    P.red_local = (const double*)(h->peer_box + b.red);
    P.rflag_local = (const unsigned long long*)(h->peer_box + b.rflag);
    return P;
}

// Allocates this rank's mailbox (after vfvm_set_halo and vfvm_set_system), writes its directory (where each neighbour's values
// land) and returns the CUDA IPC handle the host passes to the other ranks (64 bytes).
extern "C" int vfvm_peer_export(vfvm_handle* h, char handle_out[64]) {
    if (!h || !handle_out) return VFVM_ERR_ARG;
    if (h->nranks <= 1 || h->nranks > VFVM_PEER_MAX || h->n <= 0) return vfvm_fail(h, VFVM_ERR_STATE, "peer mailboxes need 2..8 ranks, a halo description and a system");
    if ((int)h->nb_ranks.size() > VFVM_PEER_MAX) return vfvm_fail(h, VFVM_ERR_STATE, "too many neighbours");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        const int R = h->nranks;
        const BoxLayout b = box_layout(R);
        const int64_t halo_doubles = (h->N - h->Nown) * h->n;
        const size_t bytes = b.halo + sizeof(double) * 2 * (size_t)std::max<int64_t>(1, halo_doubles);
        if (h->peer_box) CK(cudaFree(h->peer_box));
        CK(cudaMalloc((void**)&h->peer_box, bytes));
        CK(cudaMemset(h->peer_box, 0, bytes));
        std::vector<int64_t> dir(2 * R + 1, -1);
        for (size_t r = 0; r < h->nb_ranks.size(); r++) {
            dir[h->nb_ranks[r]] = h->recv_ptr[r];
            dir[R + h->nb_ranks[r]] = (int64_t)r;
        }
        dir[2 * R] = halo_doubles;
        CK(cudaMemcpy(h->peer_box + b.dir, dir.data(), dir.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        h->peer_counter.alloc(1);
        CK(cudaMemset(h->peer_counter.p, 0, sizeof(unsigned int)));
        cudaIpcMemHandle_t ih;
        CK(cudaIpcGetMemHandle(&ih, h->peer_box));
        static_assert(sizeof(ih) == 64, "CUDA IPC handle size");
        memcpy(handle_out, &ih, 64);
        h->peer_ok = false;
    })
    return VFVM_OK;
}

// Maps the mailboxes of all ranks (handles = nranks x 64 bytes, gathered by the host after every rank exported) and looks up
// this rank's place in each neighbour's mailbox.  On failure the NCCL transport stays in use.
extern "C" int vfvm_peer_connect(vfvm_handle* h, const char* handles) {
    if (!h || !handles || !h->peer_box) return vfvm_fail(h, VFVM_ERR_STATE, "vfvm_peer_export first");
    if (getenv("VFVM_NO_PEER")) {
        h->peer_ok = false;
        return vfvm_fail(h, VFVM_ERR_COMM, "peer mailboxes disabled by VFVM_NO_PEER");
    }
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        const int R = h->nranks;
        const BoxLayout b = box_layout(R);
        h->peer_base.assign(R, nullptr);
        h->peer_base[h->rank] = h->peer_box;
        for (int q = 0; q < R; q++) {
            if (q == h->rank) continue;
            cudaIpcMemHandle_t ih;
            memcpy(&ih, handles + (size_t)q * 64, 64);
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int k = 0; k < q; k++)
                    if (k != h->rank && h->peer_base[k]) cudaIpcCloseMemHandle(h->peer_base[k]);
                h->peer_base.clear();
                return vfvm_fail(h, VFVM_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
            h->peer_base[q] = (char*)p;
        }
        const size_t nn = h->nb_ranks.size();
        h->peer_recv_off.assign(nn, 0);
        h->peer_slot.assign(nn, 0);
        h->peer_halo_doubles.assign(nn, 0);
        for (size_t r = 0; r < nn; r++) {
            std::vector<int64_t> dir(2 * R + 1);
            CK(cudaMemcpy(dir.data(), h->peer_base[h->nb_ranks[r]] + b.dir, dir.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
            if (dir[h->rank] < 0 || dir[R + h->rank] < 0) return vfvm_fail(h, VFVM_ERR_COMM, "halo description is not symmetric between neighbours");
            h->peer_recv_off[r] = dir[h->rank];
            h->peer_slot[r] = dir[R + h->rank];
            h->peer_halo_doubles[r] = dir[2 * R];
        }
        h->peer_seq.alloc(2);  // device-resident sequence counters: [0] halo exchanges, [1] reductions
        CK(cudaMemset(h->peer_seq.p, 0, 2 * sizeof(unsigned long long)));
        if (const char* e = getenv("VFVM_PEER_TIMEOUT_MS")) h->peer_timeout_ns = std::max(1ll, atoll(e)) * 1000000ll;
        h->peer_ok = true;
    })
    return VFVM_OK;
}

extern "C" int vfvm_peer_active(vfvm_handle* h) { return h && h->peer_ok ? 1 : 0; }

// Every entry point that ran a peer kernel calls this before it returns: a wait on a peer that timed out (bit 8 of the flag word)
// must surface as VFVM_ERR_COMM -- stale halo values or partial sums would otherwise feed different Newton decisions on different
// ranks without any error.  Synchronises the stream.
int vfvm_peer_check(vfvm_handle* h) {
    if (h->nranks <= 1 || !h->peer_ok) return VFVM_OK;
    CK(cudaMemcpyAsync(h->flags_host + 3, h->flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->flags_host[3] & 256) {
        k_clear_bits<<<1, 1, 0, h->stream>>>(h->flags.p, 256);
        CK(cudaStreamSynchronize(h->stream));
        return vfvm_fail(h, VFVM_ERR_COMM, "peer exchange timed out: a neighbouring rank did not deliver its halo / partial sums (VFVM_PEER_TIMEOUT_MS)");
    }
    return VFVM_OK;
}

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
    vfvm_gather_box_free(h);
    if (h->peer_box) {
        cudaStreamSynchronize(h->stream);
        for (size_t q = 0; q < h->peer_base.size(); q++)
            if ((int)q != h->rank && h->peer_base[q]) cudaIpcCloseMemHandle(h->peer_base[q]);
        cudaFree(h->peer_box);
        h->peer_box = nullptr;
        h->peer_ok = false;
    }
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
    if (h->peer_ok) {
        const PeerArgs P = vfvm_peer_args_halo(h);
        const int64_t work = std::max<int64_t>(h->send_ptr[P.nn], P.nhalo) * h->n;
        const int grid = std::max(1, std::min(cdiv(work, 256), 148));  // every block must be resident: the last one raises the flags
        switch (h->n) {
            case 1: k_peer_halo<1><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 2: k_peer_halo<2><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 3: k_peer_halo<3><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 4: k_peer_halo<4><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 5: k_peer_halo<5><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 10: k_peer_halo<10><<<grid, 256, 0, h->stream>>>(P, x); break;
            default: throw std::string("number of species without device instantiation");
        }
        h->launches++;
        return VFVM_OK;
    }
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

// exchange arguments of a coarser AMG level: same mailboxes and neighbours as level 0, the level's own (shorter) lists
PeerArgs vfvm_peer_args_halo_level(vfvm_handle* h, const LevelHalo& c) {
    PeerArgs P = vfvm_peer_args_halo(h);
    for (int r = 0; r <= P.nn; r++) {
        P.send_ptr[r] = c.send_ptr[r];
        P.recv_ptrl[r] = c.recv_ptr[r];
    }
    P.send_idx = c.send_idx.p;
    P.Nown = c.Nown;
    P.nhalo = c.nhalo;
    return P;
}

// halo refresh of a vector of a coarser AMG level (same neighbours as level 0, shorter lists)
int vfvm_halo_exchange_level(vfvm_handle* h, LevelHalo& c, double* x) {
    if (h->nranks <= 1 || h->nb_ranks.empty()) return VFVM_OK;
    const int ns = h->n, nn = (int)h->nb_ranks.size();
    if (h->peer_ok) {
        const PeerArgs P = vfvm_peer_args_halo_level(h, c);
        const int64_t work = std::max<int64_t>(c.send_ptr[nn], c.nhalo) * ns;
        const int grid = std::max(1, std::min(cdiv(work, 256), 148));
        switch (ns) {
            case 1: k_peer_halo<1><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 2: k_peer_halo<2><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 3: k_peer_halo<3><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 4: k_peer_halo<4><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 5: k_peer_halo<5><<<grid, 256, 0, h->stream>>>(P, x); break;
            case 10: k_peer_halo<10><<<grid, 256, 0, h->stream>>>(P, x); break;
            default: throw std::string("number of species without device instantiation");
        }
        h->launches++;
        return VFVM_OK;
    }
    const int64_t nsend = c.send_ptr[nn];
    if (c.send_buf.n < (size_t)std::max<int64_t>(1, nsend) * ns) c.send_buf.alloc((size_t)std::max<int64_t>(1, nsend) * ns);
    if (nsend) {
        k_pack<<<cdiv(nsend * ns, 256), 256, 0, h->stream>>>(nsend, ns, c.send_idx.p, x, c.send_buf.p);
        h->launches++;
    }
    ncclComm_t comm = (ncclComm_t)h->nccl;
    g_nccl.GroupStart();
    for (int r = 0; r < nn; r++) {
        const int64_t s0 = c.send_ptr[r], s1 = c.send_ptr[r + 1], r0 = c.recv_ptr[r], r1 = c.recv_ptr[r + 1];
        if (s1 > s0) g_nccl.Send(c.send_buf.p + s0 * ns, (size_t)(s1 - s0) * ns, ncclFloat64, h->nb_ranks[r], comm, h->stream);
        if (r1 > r0) g_nccl.Recv(x + (c.Nown + r0) * ns, (size_t)(r1 - r0) * ns, ncclFloat64, h->nb_ranks[r], comm, h->stream);
    }
    int rc = g_nccl.GroupEnd();
    if (rc != ncclSuccess) throw std::string("NCCL halo exchange (AMG level): ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}

extern "C" int vfvm_halo_exchange(vfvm_handle* h, int which) {
    if (!h || !h->have_pattern || which < 0 || which > 3) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id or no pattern");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        vfvm_halo_exchange_ptr(h, h->vec[which].p);
        CK(cudaStreamSynchronize(h->stream));
        return vfvm_peer_check(h);
    })
}

int vfvm_comm_allreduce_sum(vfvm_handle* h, double* dev, int count) {
    if (h->nranks <= 1) return VFVM_OK;
    if (h->peer_ok && count <= VFVM_PEER_RED_W) {
        k_peer_allreduce<<<1, 32, 0, h->stream>>>(vfvm_peer_args_reduce(h), dev, count, 0);
        h->launches++;
        return VFVM_OK;
    }
    int rc = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}
// every rank contributes `count` doubles at recv + rank * count (in place: send == recv + rank * count is allowed); NCCL over NVLink.
// Used by the replicated coarse levels of the AMG hierarchy (amg.cu): one all-gather of the coarse right-hand side per cycle instead of a
// halo exchange per coarse-level SpMV.
int vfvm_comm_allgather(vfvm_handle* h, const double* send, double* recv, int64_t count) {
    if (h->nranks <= 1) {
        if (send != recv) CK(cudaMemcpyAsync(recv, send, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return VFVM_OK;
    }
    int rc = g_nccl.AllGather(send, recv, (size_t)count, ncclFloat64, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}
void vfvm_comm_group_start() {
    if (g_nccl.GroupStart) g_nccl.GroupStart();
}
void vfvm_comm_group_end() {
    if (g_nccl.GroupEnd) {
        int rc = g_nccl.GroupEnd();
        if (rc != ncclSuccess) throw std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(rc);
    }
}
// in-place all-gather of raw bytes: rank r's piece lives at buf + r * bytes_per_rank (NCCL's in-place convention)
int vfvm_comm_allgather_bytes(vfvm_handle* h, void* buf, size_t bytes_per_rank) {
    if (h->nranks <= 1 || bytes_per_rank == 0) return VFVM_OK;
    int rc = g_nccl.AllGather((const char*)buf + (size_t)h->rank * bytes_per_rank, buf, bytes_per_rank, ncclInt8, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}

// ---- gather box (replicated AMG levels) -------------------------------------------------------------------------------------
// A second IPC allocation per rank: u64 flags [2][R] at offset 0, double data [2][R][cap] at offset 256.  The IPC handles travel through
// NCCL itself (an all-gather of 64 bytes per rank), so the host API needs no extra rendezvous.  Collective: every rank calls it with the
// same capacity.  Without mapped peers (NCCL transport) nothing is allocated and vfvm_allgather_segments uses ncclAllGather.
void vfvm_gather_box_free(vfvm_handle* h) {
    if (h->gbox) {
        cudaStreamSynchronize(h->stream);
        for (size_t q = 0; q < h->gbox_base.size(); q++)
            if ((int)q != h->rank && h->gbox_base[q]) cudaIpcCloseMemHandle(h->gbox_base[q]);
        cudaFree(h->gbox);
    }
    h->gbox = nullptr;
    h->gbox_base.clear();
    h->gbox_cap = 0;
    h->gbox_ok = false;
}
int vfvm_gather_box_create(vfvm_handle* h, int64_t cap) {
    if (h->nranks <= 1 || !h->peer_ok || getenv("VFVM_NO_GATHER_BOX")) return VFVM_OK;
    if (h->gbox_ok && h->gbox_cap >= cap) return VFVM_OK;
    vfvm_gather_box_free(h);
    const int R = h->nranks;
    const size_t bytes = 256 + sizeof(double) * 2 * (size_t)R * (size_t)cap;
    CK(cudaMalloc((void**)&h->gbox, bytes));
    CK(cudaMemset(h->gbox, 0, bytes));
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, h->gbox));
    DevBuf<char> hb;
    hb.alloc((size_t)R * 64);
    CK(cudaMemcpyAsync(hb.p + (size_t)h->rank * 64, &ih, 64, cudaMemcpyHostToDevice, h->stream));
    vfvm_comm_allgather_bytes(h, hb.p, 64);
    std::vector<char> all = hb.to_host(h->stream);
    h->gbox_base.assign(R, nullptr);
    h->gbox_base[h->rank] = h->gbox;
    bool ok = true;
    for (int q = 0; q < R && ok; q++) {
        if (q == h->rank) continue;
        cudaIpcMemHandle_t qh;
        memcpy(&qh, all.data() + (size_t)q * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, qh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        } else {
            h->gbox_base[q] = (char*)p;
        }
    }
    // every rank must take the same path: agree on the outcome
    DevBuf<double> flag;
    flag.alloc(1);
    const double mine = ok ? 0.0 : 1.0;
    CK(cudaMemcpyAsync(flag.p, &mine, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    vfvm_comm_allreduce_max(h, flag.p, 1);
    double any = 0.0;
    CK(cudaMemcpyAsync(&any, flag.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (any != 0.0) {
        vfvm_gather_box_free(h);
        return VFVM_OK;  // NCCL all-gather instead
    }
    h->gbox_seq.alloc(1);
    h->gbox_count.alloc(1);
    CK(cudaMemsetAsync(h->gbox_seq.p, 0, sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->gbox_count.p, 0, sizeof(unsigned int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->gbox_cap = cap;
    h->gbox_ok = true;
    return VFVM_OK;
}
int vfvm_allgather_segments(vfvm_handle* h, double* vec, int64_t cap) {
    if (h->nranks <= 1) return VFVM_OK;
    if (h->gbox_ok && cap <= h->gbox_cap && (cap & 1) == 0) {
        GatherArgs G;
        memset(&G, 0, sizeof(G));
        G.nranks = h->nranks;
        G.rank = h->rank;
        G.cap = cap;
        G.seq_ctr = h->gbox_seq.p;
        G.count = h->gbox_count.p;
        for (int q = 0; q < h->nranks; q++) {
            G.dst[q] = (double*)(h->gbox_base[q] + 256);
            G.flag_dst[q] = (unsigned long long*)h->gbox_base[q];
        }
        G.box_local = (const double*)(h->gbox + 256);
        G.flag_local = (const unsigned long long*)h->gbox;
        G.err = h->flags.p;
        G.timeout_ns = h->peer_timeout_ns;
        const int grid = std::max(1, std::min(cdiv(cap * (h->nranks - 1), 2 * 256), 148));  // one resident wave
        k_peer_allgather<<<grid, 256, 0, h->stream>>>(G, vec);
        h->launches++;
        return VFVM_OK;
    }
    return vfvm_comm_allgather(h, vec + (int64_t)h->rank * cap, vec, cap);
}
int vfvm_comm_allreduce_max(vfvm_handle* h, double* dev, int count) {
    if (h->nranks <= 1) return VFVM_OK;
    if (h->peer_ok && count <= VFVM_PEER_RED_W) {
        k_peer_allreduce<<<1, 32, 0, h->stream>>>(vfvm_peer_args_reduce(h), dev, count, 1);
        h->launches++;
        return VFVM_OK;
    }
    int rc = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, (ncclComm_t)h->nccl, h->stream);
    if (rc != ncclSuccess) throw std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc);
    return VFVM_OK;
}
