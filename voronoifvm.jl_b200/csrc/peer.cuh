// Peer-memory exchange between the ranks of one NVSwitch domain (one process per GPU, CUDA IPC).
//
// Every rank owns one "mailbox" allocation that all peers map.  Producers STORE into the consumer's mailbox over NVLink
// and then raise a sequence flag there; consumers spin on their LOCAL flag and read their LOCAL mailbox, so no kernel ever
// waits on a remote load.  Two uses:
//   * halo of the SpMV input: the SpMV kernel itself pushes this rank's boundary values into its neighbours' mailboxes
//     (first thing it does), multiplies all rows whose columns are owned, and only the warps that reach a halo column wait
//     for the neighbours' flags -- one kernel per SpMV, the transfer hides behind the interior rows;
//   * Krylov inner products / norms: the single-block finalize kernel stores its partial sums into every peer's
//     reduction box, waits for the peers' sums and adds them in rank order (identical bits on every rank).
// Both are double-buffered by the parity of their sequence number.  A producer can be at most one exchange ahead of a
// consumer (its exchange k+1 needs the consumer's push k+1, which is stream-ordered after the consumer's read of k),
// so exchange k+2 can never overwrite data of exchange k that is still being read.
// NCCL remains the transport when the peers cannot be mapped (vfvm_peer_connect fails) or VFVM_NO_PEER is set.
#pragma once
#include <cstdint>

#define VFVM_PEER_MAX 8      // ranks of one NVSwitch domain
#define VFVM_PEER_RED_W 32   // doubles per rank in a reduction box (GMRES multi-dot needs restart+1)
#define VFVM_PEER_SPIN_MAX (1ll << 23)  // bounded spinning: a lost peer becomes VFVM_ERR_COMM instead of a hung GPU

// kernel argument of one exchange (the parity of `seq` is already selected by the host)
struct PeerArgs {
    int nn, nranks, rank, ns;
    unsigned long long seq;
    int64_t send_ptr[VFVM_PEER_MAX + 1];  // my send list, grouped by neighbour slot
    const int32_t* send_idx;
    int64_t Nown, nhalo;
    // generic exchange of a coarser level through the level-0 mailbox: neighbour r's values start at recv_ptr0[r] in the mailbox and
    // at recv_ptrl[r] in the local halo order (equal on level 0)
    int64_t recv_ptr0[VFVM_PEER_MAX + 1], recv_ptrl[VFVM_PEER_MAX + 1];
    // remote (peer-mapped) addresses per neighbour slot: where MY boundary values / my flag go
    double* halo_dst[VFVM_PEER_MAX];
    unsigned long long* hflag_dst[VFVM_PEER_MAX];
    const double* halo_local;                // my mailbox: halo values in local halo order
    const unsigned long long* hflag_local;   // [slot]
    unsigned int* push_count;                // grid-wide completion counter of the push phase
    // reductions, per rank q (including myself): q's box row for me
    double* red_dst[VFVM_PEER_MAX];
    unsigned long long* rflag_dst[VFVM_PEER_MAX];
    const double* red_local;                 // [rank][VFVM_PEER_RED_W]
    const unsigned long long* rflag_local;   // [rank]
    int32_t* err;                            // device flag word: bit 8 = peer timeout
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long peer_ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double peer_ld_data(const double* p) {  // mailbox data is written by peers: bypass L1
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// Push phase of a halo exchange, executed by every block of the calling kernel: this rank's boundary values of x go straight
// into the neighbours' mailboxes; the block that finishes last raises this rank's flag at every neighbour.
template <int NS>
__device__ __forceinline__ void peer_push(const PeerArgs& P, const double* __restrict__ x) {
    const int64_t total = P.send_ptr[P.nn] * NS;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / NS;
        const int s = (int)(i - q * NS);
        int r = 0;
        while (q >= P.send_ptr[r + 1]) r++;
        P.halo_dst[r][(q - P.send_ptr[r]) * NS + s] = x[(int64_t)P.send_idx[q] * NS + s];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(P.push_count, 1u);
        if (prev == gridDim.x - 1) {
            *P.push_count = 0;  // every block has arrived: ready for the next launch
            __threadfence_system();
            for (int r = 0; r < P.nn; r++) peer_st_flag(P.hflag_dst[r], P.seq);
        }
    }
}

// mailbox position of local halo node c (level 0: the identity)
__device__ __forceinline__ int64_t peer_halo_pos(const PeerArgs& P, int64_t c) {
    int r = 0;
    while (c >= P.recv_ptrl[r + 1]) r++;
    return P.recv_ptr0[r] + (c - P.recv_ptrl[r]);
}

// spin until *flag >= seq (bounded); returns false on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long* flag, unsigned long long seq, int32_t* err) {
    long long spins = 0;
    if (*(volatile int32_t*)err & 256) return false;  // a peer is already lost: do not wait again
    while (peer_ld_flag(flag) < seq) {
        if (++spins > VFVM_PEER_SPIN_MAX) {
            atomicOr(err, 256);
            return false;
        }
        __nanosleep(20);
    }
    return true;
}
#endif
