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

// Kernel argument of one exchange.  The sequence number of the exchange lives in DEVICE memory (`seq_ctr`, one counter for the halo
// exchanges and one for the reductions): every kernel that takes part reads it at its start, uses seq = counter + 1, and the block
// that finishes the push phase last (by then every block has read the counter) stores seq back.  Nothing in the argument depends on
// the sequence number -- the two parities' buffers are addressed as base + parity * stride inside the kernel -- so a captured CUDA
// graph of a whole Krylov iteration (SpMV with its halo push, the AMG cycle with one exchange per level SpMV, the reductions) can be
// replayed any number of times.
struct PeerArgs {
    int nn, nranks, rank, ns;
    unsigned long long* seq_ctr;
    int64_t send_ptr[VFVM_PEER_MAX + 1];  // my send list, grouped by neighbour slot
    const int32_t* send_idx;
    int64_t Nown, nhalo;
    // generic exchange of a coarser level through the level-0 mailbox: neighbour r's values start at recv_ptr0[r] in the mailbox and
    // at recv_ptrl[r] in the local halo order (equal on level 0)
    int64_t recv_ptr0[VFVM_PEER_MAX + 1], recv_ptrl[VFVM_PEER_MAX + 1];
    // remote (peer-mapped) addresses per neighbour slot, parity 0: where MY boundary values / my flag go
    double* halo_dst[VFVM_PEER_MAX];
    int64_t halo_dst_stride[VFVM_PEER_MAX];  // doubles between the two parities in that neighbour's mailbox
    unsigned long long* hflag_dst[VFVM_PEER_MAX];  // parity stride: nranks
    const double* halo_local;                // my mailbox, parity 0: halo values in local halo order
    int64_t halo_local_stride;
    const unsigned long long* hflag_local;   // [parity][slot]
    unsigned int* push_count;                // grid-wide completion counter of the push phase
    // reductions, per rank q (including myself), parity 0: q's box row for me (parity strides: nranks * VFVM_PEER_RED_W / nranks)
    double* red_dst[VFVM_PEER_MAX];
    unsigned long long* rflag_dst[VFVM_PEER_MAX];
    const double* red_local;                 // [parity][rank][VFVM_PEER_RED_W]
    const unsigned long long* rflag_local;   // [parity][rank]
    int32_t* err;                            // device flag word: bit 8 = peer timeout
    long long timeout_ns;                    // bound of a wait on a peer (a lost peer becomes VFVM_ERR_COMM instead of a hung GPU)
};

// Kernel argument of an all-gather through the gather boxes (replicated AMG levels): every rank stores its segment into every peer's
// box [parity][rank][cap] and raises flag [parity][rank] there; the consumer waits on its local flags and copies the peers' segments
// out of its local box.  Same sequence-counter scheme as above.
struct GatherArgs {
    int nranks, rank;
    int64_t cap;                              // doubles per segment
    unsigned long long* seq_ctr;
    unsigned int* count;                      // grid-wide completion counter of the push phase
    double* dst[VFVM_PEER_MAX];               // rank q's box data, parity 0, segment 0 (peer-mapped)
    unsigned long long* flag_dst[VFVM_PEER_MAX];
    const double* box_local;
    const unsigned long long* flag_local;     // [parity][rank]
    int32_t* err;
    long long timeout_ns;
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
__device__ __forceinline__ long long peer_now_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// sequence number of the exchange this kernel performs (every thread reads the same value: the counter is only advanced once every
// block of the kernel has read it)
__device__ __forceinline__ unsigned long long peer_seq(const PeerArgs& P) { return *(volatile unsigned long long*)P.seq_ctr + 1ull; }

// Push phase of a halo exchange, executed by every block of the calling kernel: this rank's boundary values of x go straight
// into the neighbours' mailboxes; the block that finishes last raises this rank's flag at every neighbour and advances the counter.
template <int NS>
__device__ __forceinline__ void peer_push(const PeerArgs& P, unsigned long long seq, const double* __restrict__ x) {
    const int64_t total = P.send_ptr[P.nn] * NS;
    const int par = (int)(seq & 1ull);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / NS;
        const int s = (int)(i - q * NS);
        int r = 0;
        while (q >= P.send_ptr[r + 1]) r++;
        P.halo_dst[r][par * P.halo_dst_stride[r] + (q - P.send_ptr[r]) * NS + s] = x[(int64_t)P.send_idx[q] * NS + s];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(P.push_count, 1u);
        if (prev == gridDim.x - 1) {
            *P.push_count = 0;  // every block has arrived (and has read the sequence counter): ready for the next launch
            *P.seq_ctr = seq;
            __threadfence_system();
            for (int r = 0; r < P.nn; r++) peer_st_flag(P.hflag_dst[r] + par * P.nranks, seq);
        }
    }
}

// mailbox position of local halo node c (level 0: the identity)
__device__ __forceinline__ int64_t peer_halo_pos(const PeerArgs& P, int64_t c) {
    int r = 0;
    while (c >= P.recv_ptrl[r + 1]) r++;
    return P.recv_ptr0[r] + (c - P.recv_ptrl[r]);
}
__device__ __forceinline__ const double* peer_halo_local(const PeerArgs& P, unsigned long long seq) { return P.halo_local + (int64_t)(seq & 1ull) * P.halo_local_stride; }
__device__ __forceinline__ const unsigned long long* peer_hflag_local(const PeerArgs& P, unsigned long long seq) { return P.hflag_local + (seq & 1ull) * P.nranks; }

// spin until *flag >= seq, bounded in TIME (rank skew of seconds is legitimate, e.g. a host-side hierarchy build on one rank);
// returns false on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long* flag, unsigned long long seq, int32_t* err, long long timeout_ns) {
    if (peer_ld_flag(flag) >= seq) return true;
    if (*(volatile int32_t*)err & 256) return false;  // a peer is already lost: do not wait again
    const long long t0 = peer_now_ns();
    int spins = 0;
    while (peer_ld_flag(flag) < seq) {
        if ((++spins & 63) == 0 && peer_now_ns() - t0 > timeout_ns) {
            atomicOr(err, 256);
            return false;
        }
        __nanosleep(20);
    }
    return true;
}
// single-block all-reduce step shared by the reduction kernels: thread t < nranks delivers `count` values to rank t and waits for
// rank t's values; afterwards red (parity-selected, [rank][VFVM_PEER_RED_W]) holds every rank's contribution
__device__ __forceinline__ const double* peer_reduce_exchange(const PeerArgs& P, unsigned long long seq, const double* mine, int count) {
    const int t = threadIdx.x, par = (int)(seq & 1ull);
    if (t < P.nranks) {
        double* dst = P.red_dst[t] + (int64_t)par * P.nranks * VFVM_PEER_RED_W;
        for (int v = 0; v < count; v++) dst[v] = mine[v];
        __threadfence_system();
        peer_st_flag(P.rflag_dst[t] + par * P.nranks, seq);
        peer_wait(P.rflag_local + par * P.nranks + t, seq, P.err, P.timeout_ns);
    }
    __syncthreads();
    if (t == 0) *P.seq_ctr = seq;  // single-block kernel: every thread has read the counter before the barrier above
    return P.red_local + (int64_t)par * P.nranks * VFVM_PEER_RED_W;
}
#endif
