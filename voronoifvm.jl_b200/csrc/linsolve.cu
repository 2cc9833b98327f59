// K8-K11: sparse linear solve of a Newton step, _solve_linear! (src/vfvm_linsolve.jl:6-61), plus the Newton update.
//
// Replaces LinearSolve.jl/Krylov.jl/ILUZero.jl (third party, not under /root/reference) by
//   * SpMV on the DBSR planes over the same row tiles as the assembly kernel (one thread per off-diagonal block,
//     fixed-order row reduction in shared memory, diagonal block applied by the row thread), with the Krylov inner
//     products fused into its row phase;
//   * BiCGStab / CG whose scalars (rho, alpha, omega, ...) live in device memory -- no host round trip inside an
//     iteration except one pinned-memory read of ||r||^2 for the stopping test;
//   * Jacobi / node-block Jacobi preconditioners (ILU0: ilu0.cu).
// All reductions are two-stage with a fixed order => bitwise reproducible.  With more than one rank the partial sums
// are combined by ncclAllReduce on the same stream and the SpMV input gets a halo refresh first (comm.cu).
#include <algorithm>

#include "bulk.cuh"
#include "vfvm_internal.h"

#define LS_THREADS 256
#define LS_RMAX 256

// device scalar slots
enum { S_RHO = 0, S_RHO_OLD, S_ALPHA, S_OMEGA, S_BETA, S_RV, S_RESTART, S_RHAT2, S_RR, S_BB, S_PAP, S_RZ, S_RZ_OLD, S_TMP0, S_TMP1, S_COUNT = 16 };

int vfvm_comm_allreduce_sum(vfvm_handle* h, double* dev, int count);
int vfvm_comm_allreduce_max(vfvm_handle* h, double* dev, int count);
int vfvm_halo_exchange_ptr(vfvm_handle* h, double* x);
void vfvm_ilu0_setup(vfvm_handle* h);
void vfvm_ilu0_apply(vfvm_handle* h, const double* in, double* out);

namespace {


__device__ __forceinline__ double block_sum(double v, double* sh) {
    // fixed-order tree reduction over LS_THREADS threads
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += sh[i];
    return r;  // valid in thread 0
}
__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) r = fmax(r, sh[i]);
    return r;
}

// y = A x on the DBSR planes in SELL-32 order: a warp owns a slice of 32 rows, one lane per row; colidx / value loads are
// coalesced, x[col] gathers coalesce on structured numberings, the row sum lives in registers (column order, deterministic).
// Optional fused inner products (y,w) and (y,y) for the Krylov recurrences.
// PEER (several ranks, peer.cuh): the kernel first pushes this rank's boundary values of x into the neighbours' mailboxes and
// raises its flag there; columns >= Nown are read from the local mailbox, and only a warp that reaches such a column waits
// for the neighbours' flags -- halo exchange and SpMV are one kernel, the transfer hides behind the interior rows.
template <int NS, bool DIAGMASK, bool PEER, class VT = double>
__global__ void __launch_bounds__(LS_THREADS) k_spmv(const SpmvArgs a, const PeerArgs P) {
    __shared__ double red[32];
    const VT* __restrict__ vals;  // fp64 planes, or their fp32 copy (preconditioner-internal products)
    if constexpr (sizeof(VT) == 4) vals = (const VT*)a.offval32;
    else vals = (const VT*)a.offval;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const int64_t nnz = a.nnz_sell;
    double d_yw = 0.0, d_yy = 0.0;
    bool halo_ready = false;  // warp-uniform
    unsigned long long seq = 0;
    const double* __restrict__ hbox = nullptr;
    if constexpr (PEER) {
        seq = peer_seq(P);
        hbox = peer_halo_local(P, seq);
        peer_push<NS>(P, seq, a.x);
    }
    for (int g = blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double acc[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) acc[i] = 0.0;
        constexpr int BATCH = NS == 1 ? 8 : (NS <= 3 ? 4 : 2);
        for (int j0 = 0; j0 < w; j0 += BATCH) {
            int Lc[BATCH];
            double v1[BATCH];  // NS == 1: the matrix value travels with the index load
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                const bool ok = j0 + b < w;
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
                Lc[b] = ok ? a.colidx[e] : (int)r;
                v1[b] = (NS == 1 && DIAGMASK && ok) ? (double)vals[e] : 0.0;
            }
            double xl[BATCH][NS];
            if constexpr (PEER) {
                bool need = false;
#pragma unroll
                for (int b = 0; b < BATCH; b++) need |= Lc[b] >= a.Nown;
                if (!halo_ready && __any_sync(0xffffffffu, need)) {
                    if (lane < P.nn) peer_wait(peer_hflag_local(P, seq) + lane, seq, P.err, P.timeout_ns);
                    __syncwarp();
                    halo_ready = true;
                }
#pragma unroll
                for (int b = 0; b < BATCH; b++)
#pragma unroll
                    for (int jj = 0; jj < NS; jj++)
                        xl[b][jj] = Lc[b] >= a.Nown ? peer_ld_data(hbox + peer_halo_pos(P, Lc[b] - a.Nown) * NS + jj) : a.x[(int64_t)Lc[b] * NS + jj];
            } else {
#pragma unroll
                for (int b = 0; b < BATCH; b++)
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) xl[b][jj] = a.x[(int64_t)Lc[b] * NS + jj];
            }
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (j0 + b >= w) break;
                if constexpr (NS == 1 && DIAGMASK) {
                    acc[0] += v1[b] * xl[b][0];
                } else if constexpr (DIAGMASK) {  // species-decoupled planes: plane i <-> (i,i)
                    const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
#pragma unroll
                    for (int i = 0; i < NS; i++) acc[i] += (double)vals[(int64_t)i * nnz + e] * xl[b][i];
                } else {
                    const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
#pragma unroll
                    for (int i = 0; i < NS; i++)
#pragma unroll
                        for (int jj = 0; jj < NS; jj++) {
                            const int p = a.idxF[i * NS + jj];
                            if (p >= 0) acc[i] += (double)vals[(int64_t)p * nnz + e] * xl[b][jj];
                        }
                }
            }
        }
        if (valid) {
            double xr[NS];
#pragma unroll
            for (int jj = 0; jj < NS; jj++) xr[jj] = a.x[r * NS + jj];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double s = acc[i];
                if constexpr (DIAGMASK) {
                    s += a.diagval[(int64_t)i * a.Nown + r] * xr[i];
                } else {
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) {
                        const int p = a.idxD[i * NS + jj];
                        if (p >= 0) s += a.diagval[(int64_t)p * a.Nown + r] * xr[jj];
                    }
                }
                a.y[r * NS + i] = s;
                if (a.w) {
                    d_yw += s * a.w[r * NS + i];
                    d_yy += s * s;
                }
            }
        }
    }
    if (a.w) {
        const double s1 = block_sum(d_yw, red);
        const double s2 = block_sum(d_yy, red);
        if (threadIdx.x == 0) {
            a.part[blockIdx.x] = s1;
            a.part[gridDim.x + blockIdx.x] = s2;
        }
    }
}


// The same product with the index and value planes streamed through shared memory by asynchronous bulk copies (1-D TMA).
// Motivation (ncu launch list of a cfg4 Newton step, profiles/r2_*): with n x n coupled species the register-staged kernel above holds
// BATCH x (n + planes) doubles per lane, runs at 16 warps per SM and reaches 2.7 TB/s -- each batch waits for its column indices and
// then for its gathers and values, and nothing else is in flight meanwhile.  Here every warp owns a ring of BULK_STAGES stages in
// shared memory; lane 0 arms the stage's mbarrier and issues one bulk copy per plane for a chunk of BULK_CH entries per row
// (32 rows x BULK_CH entries: 128 B-aligned contiguous pieces of every plane, because a slice is stored entry-major), so
// BULK_STAGES chunks per warp are in flight at any time without occupying registers, across slice boundaries.  The lanes read the
// chunk from shared memory (consecutive lanes, no bank conflicts), gather x and accumulate exactly as above -- the summation order
// per row is unchanged, so the result is bitwise the same.
#define BULK_CH 4
#define BULK_STAGES 3
#define BULK_THREADS 128
template <int NS, bool DIAGMASK, bool PEER, class VT = double>
__global__ void __launch_bounds__(BULK_THREADS) k_spmv_bulk(const SpmvArgs a, const PeerArgs P, const int cF) {
    constexpr int VB = (int)sizeof(VT);
    const VT* __restrict__ vals;
    if constexpr (VB == 4) vals = (const VT*)a.offval32;
    else vals = (const VT*)a.offval;
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t bars[(BULK_THREADS / 32) * BULK_STAGES];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const int64_t nnz = a.nnz_sell;
    const int stage_bytes = BULK_CH * 32 * 4 + cF * BULK_CH * 32 * VB;
    unsigned char* wbase = bulk_smem + (size_t)wib * BULK_STAGES * stage_bytes;
    uint64_t* wbar = bars + wib * BULK_STAGES;
    if (lane == 0)
        for (int s = 0; s < BULK_STAGES; s++) mbar_init(wbar + s, 1);
    mbar_fence_init();
    __syncthreads();
    double d_yw = 0.0, d_yy = 0.0;
    bool halo_ready = false;  // warp-uniform
    unsigned long long seq = 0;
    const double* __restrict__ hbox = nullptr;
    if constexpr (PEER) {
        seq = peer_seq(P);
        hbox = peer_halo_local(P, seq);
        peer_push<NS>(P, seq, a.x);
    }
    const int g0 = blockIdx.x * wpb + wib;
    const uint64_t stream_policy = l2_policy_evict_first();  // the planes are read once: keep x (gathered ~15 times per entry) in L2 instead
    // producer cursor (lane 0): next chunk to request
    int pg = g0, pj0 = 0, pbase = 0, pw = 0;
    auto producer_seek = [&]() {  // skip slices without entries
        while (pg < a.nslices) {
            pbase = a.sell_ptr[pg];
            pw = (a.sell_ptr[pg + 1] - pbase) >> 5;
            if (pw > 0) break;
            pg += nwarps;
        }
    };
    auto issue = [&](int s) {
        if (pg >= a.nslices) return;
        const int nent = min(BULK_CH, pw - pj0) * 32;
        const int64_t e0 = (int64_t)pbase + (int64_t)pj0 * 32;
        unsigned char* st = wbase + (size_t)s * stage_bytes;
        mbar_expect_tx(wbar + s, (uint32_t)(nent * 4 + cF * nent * VB));
        bulk_g2s_hint(st, a.colidx + e0, (uint32_t)(nent * 4), wbar + s, stream_policy);
        for (int p = 0; p < cF; p++)
            bulk_g2s_hint(st + BULK_CH * 32 * 4 + (size_t)p * BULK_CH * 32 * VB, vals + (int64_t)p * nnz + e0, (uint32_t)(nent * VB), wbar + s, stream_policy);
        pj0 += BULK_CH;
        if (pj0 >= pw) {
            pg += nwarps;
            pj0 = 0;
            producer_seek();
        }
    };
    if (lane == 0) {
        producer_seek();
        for (int s = 0; s < BULK_STAGES; s++) issue(s);
    }
    int stage = 0;
    uint32_t phases = 0;  // bit s: parity the consumer waits for on stage s
    for (int g = g0; g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double acc[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) acc[i] = 0.0;
        for (int j0 = 0; j0 < w; j0 += BULK_CH) {
            mbar_wait(wbar + stage, (phases >> stage) & 1u);
            const unsigned char* st = wbase + (size_t)stage * stage_bytes;
            const int32_t* __restrict__ s_col = (const int32_t*)st;
            const VT* __restrict__ s_val = (const VT*)(st + BULK_CH * 32 * 4);
            int Lc[BULK_CH];
#pragma unroll
            for (int b = 0; b < BULK_CH; b++) Lc[b] = (j0 + b < w) ? s_col[b * 32 + lane] : (int)r;
            double xl[BULK_CH][NS];
            if constexpr (PEER) {
                bool need = false;
#pragma unroll
                for (int b = 0; b < BULK_CH; b++) need |= Lc[b] >= a.Nown;
                if (!halo_ready && __any_sync(0xffffffffu, need)) {
                    if (lane < P.nn) peer_wait(peer_hflag_local(P, seq) + lane, seq, P.err, P.timeout_ns);
                    __syncwarp();
                    halo_ready = true;
                }
#pragma unroll
                for (int b = 0; b < BULK_CH; b++)
#pragma unroll
                    for (int jj = 0; jj < NS; jj++)
                        xl[b][jj] = Lc[b] >= a.Nown ? peer_ld_data(hbox + peer_halo_pos(P, Lc[b] - a.Nown) * NS + jj) : a.x[(int64_t)Lc[b] * NS + jj];
            } else {
#pragma unroll
                for (int b = 0; b < BULK_CH; b++)
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) xl[b][jj] = a.x[(int64_t)Lc[b] * NS + jj];
            }
#pragma unroll
            for (int b = 0; b < BULK_CH; b++) {
                if (j0 + b >= w) break;
                if constexpr (DIAGMASK) {
#pragma unroll
                    for (int i = 0; i < NS; i++) acc[i] += (double)s_val[i * BULK_CH * 32 + b * 32 + lane] * xl[b][i];
                } else {
#pragma unroll
                    for (int i = 0; i < NS; i++)
#pragma unroll
                        for (int jj = 0; jj < NS; jj++) {
                            const int p = a.idxF[i * NS + jj];
                            if (p >= 0) acc[i] += (double)s_val[p * BULK_CH * 32 + b * 32 + lane] * xl[b][jj];
                        }
                }
            }
            __syncwarp();  // every lane has read the stage
            if (lane == 0) {
                fence_proxy_async_smem();
                issue(stage);
            }
            phases ^= 1u << stage;
            stage = stage + 1 == BULK_STAGES ? 0 : stage + 1;
        }
        if (valid) {
            double xr[NS];
#pragma unroll
            for (int jj = 0; jj < NS; jj++) xr[jj] = a.x[r * NS + jj];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double s = acc[i];
                if constexpr (DIAGMASK) {
                    s += a.diagval[(int64_t)i * a.Nown + r] * xr[i];
                } else {
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) {
                        const int p = a.idxD[i * NS + jj];
                        if (p >= 0) s += a.diagval[(int64_t)p * a.Nown + r] * xr[jj];
                    }
                }
                a.y[r * NS + i] = s;
                if (a.w) {
                    d_yw += s * a.w[r * NS + i];
                    d_yy += s * s;
                }
            }
        }
    }
    if (a.w) {
        const double s1 = block_sum(d_yw, red);
        const double s2 = block_sum(d_yy, red);
        if (threadIdx.x == 0) {
            a.part[blockIdx.x] = s1;
            a.part[gridDim.x + blockIdx.x] = s2;
        }
    }
}

// sums `nparts` partials for each of `nvals` values into out[0..nvals) in fixed order (single block)
__global__ void k_finalize(const double* __restrict__ part, int nparts, int nvals, double* __restrict__ out) {
    __shared__ double red[32];
    for (int v = 0; v < nvals; v++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[(int64_t)v * nparts + i];
        const double r = block_sum(s, red);
        if (threadIdx.x == 0) out[v] = r;
        __syncthreads();
    }
}
__global__ void k_finalize_max(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s = fmax(s, part[i]);
    const double r = block_max(s, red);
    if (threadIdx.x == 0) out[0] = r;
}

// ---- scalar recurrences on the device; fused with the final reduction when there is a single rank ------------------
enum { OP_NONE = 0, OP_BICG_INIT, OP_BICG_ALPHA, OP_BICG_OMEGA, OP_BICG_NEXT, OP_CG_INIT, OP_CG_ALPHA, OP_CG_NEXT };
__device__ void scalar_update(int op, double* __restrict__ sc, int32_t* __restrict__ flags) {
    switch (op) {
        case OP_BICG_INIT:  // TMP0 = (r,r)
            sc[S_RHO] = sc[S_TMP0];
            sc[S_RR] = sc[S_TMP0];
            sc[S_BB] = sc[S_TMP0];
            sc[S_ALPHA] = 1.0;
            sc[S_OMEGA] = 1.0;
            sc[S_BETA] = 0.0;
            sc[S_RESTART] = 0.0;
            sc[S_RHAT2] = sc[S_TMP0];
            break;
        case OP_BICG_ALPHA:  // TMP0 = (rhat, v)
            sc[S_RV] = sc[S_TMP0];
            sc[S_ALPHA] = sc[S_RHO] / sc[S_RV];
            if (!(fabs(sc[S_RV]) > 0.0) || sc[S_ALPHA] != sc[S_ALPHA]) atomicOr(flags, 2);
            break;
        case OP_BICG_OMEGA:  // TMP0 = (t,s), TMP1 = (t,t)
            sc[S_OMEGA] = (sc[S_TMP1] > 0.0) ? sc[S_TMP0] / sc[S_TMP1] : 0.0;
            break;
        case OP_BICG_NEXT: {  // TMP0 = (rhat, r_new), TMP1 = (r_new, r_new): beta for the next iteration
            const double rho_new = sc[S_TMP0], rr = sc[S_TMP1];
            sc[S_RR] = rr;
            // near-breakdown: the shadow residual has become (numerically) orthogonal to the residual, rho would underflow
            // within a few iterations.  Restart with rhat = r (done by the next k_bicg_p): rho = (r,r), beta = 0.
            if (rho_new * rho_new < 1.0e-16 * rr * sc[S_RHAT2] || !(fabs(sc[S_OMEGA]) > 0.0)) {
                sc[S_RESTART] = 1.0;
                sc[S_RHAT2] = rr;
                sc[S_RHO] = rr;
                sc[S_BETA] = 0.0;
            } else {
                sc[S_RESTART] = 0.0;
                sc[S_BETA] = (rho_new / sc[S_RHO]) * (sc[S_ALPHA] / sc[S_OMEGA]);
                sc[S_RHO] = rho_new;
            }
            break;
        }
        case OP_CG_INIT:  // TMP0 = (z,r), TMP1 = (r,r)
            sc[S_RZ] = sc[S_TMP0];
            sc[S_RR] = sc[S_TMP1];
            sc[S_BB] = sc[S_TMP1];
            sc[S_BETA] = 0.0;
            break;
        case OP_CG_ALPHA:  // TMP0 = (Ap, p)
            sc[S_ALPHA] = sc[S_RZ] / sc[S_TMP0];
            if (!(fabs(sc[S_TMP0]) > 0.0)) atomicOr(flags, 2);
            break;
        case OP_CG_NEXT:  // TMP0 = new (z,r), TMP1 = (r,r)
            sc[S_BETA] = sc[S_TMP0] / sc[S_RZ];
            sc[S_RZ] = sc[S_TMP0];
            sc[S_RR] = sc[S_TMP1];
            break;
        default: break;
    }
}
__global__ void k_scalar(int op, double* __restrict__ sc, int32_t* __restrict__ flags) { scalar_update(op, sc, flags); }

// sums `nparts` partials for each of `nvals` values into sc[S_TMP0 + v] in fixed order (single block), then applies `op`
__global__ void k_finalize_op(const double* __restrict__ part, int nparts, int nvals, double* __restrict__ sc, int op, int32_t* __restrict__ flags) {
    __shared__ double red[32];
    for (int v = 0; v < nvals; v++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[(int64_t)v * nparts + i];
        const double r = block_sum(s, red);
        if (threadIdx.x == 0) sc[S_TMP0 + v] = r;
        __syncthreads();
    }
    if (threadIdx.x == 0 && op != OP_NONE) scalar_update(op, sc, flags);
}

// several ranks with peer mailboxes: local sums -> every rank's reduction box -> wait -> sum over ranks in rank order -> `op`,
// all in this one single-block kernel (no ncclAllReduce, no separate scalar kernel)
__global__ void k_finalize_op_peer(const double* __restrict__ part, int nparts, int nvals, double* __restrict__ sc, int op, int32_t* __restrict__ flags,
                                   const PeerArgs P) {
    __shared__ double red[32];
    __shared__ double mine[VFVM_PEER_RED_W];
    for (int v = 0; v < nvals; v++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[(int64_t)v * nparts + i];
        const double r = block_sum(s, red);
        if (threadIdx.x == 0) mine[v] = r;
        __syncthreads();
    }
    const int t = threadIdx.x;
    const unsigned long long seq = peer_seq(P);
    const double* red_all = peer_reduce_exchange(P, seq, mine, nvals);
    if (t < nvals) {
        double acc = peer_ld_data(red_all + t);
        for (int q = 1; q < P.nranks; q++) acc += peer_ld_data(red_all + (size_t)q * VFVM_PEER_RED_W + t);
        sc[S_TMP0 + t] = acc;
    }
    __syncthreads();
    if (t == 0 && op != OP_NONE) scalar_update(op, sc, flags);
}

// ---- fused vector kernels (grid-stride over n*Nown entries; partial sums per block) -------------------------
// `dinv` = reciprocal point diagonal (Jacobi) or null (identity): the preconditioned vector is produced in the same pass
// BiCGStab: p = r + beta (p - omega v) ; phat = M^-1 p
__global__ void k_bicg_p(int64_t n, const double* __restrict__ sc, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ p,
                         const double* __restrict__ dinv, double* __restrict__ phat, double* __restrict__ rhat) {
    const double beta = sc[S_BETA], omega = sc[S_OMEGA];
    const bool restart = sc[S_RESTART] != 0.0;  // near-breakdown detected by the last scalar update: new shadow residual
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (restart) rhat[i] = r[i];
        const double pi = restart ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        if (phat) phat[i] = dinv ? pi * dinv[i] : pi;
    }
}
// s = r - alpha v ; shat = M^-1 s
__global__ void k_bicg_s(int64_t n, const double* __restrict__ sc, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ s,
                         const double* __restrict__ dinv, double* __restrict__ shat) {
    const double alpha = sc[S_ALPHA];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double si = r[i] - alpha * v[i];
        s[i] = si;
        if (shat) shat[i] = dinv ? si * dinv[i] : si;
    }
}
// x += alpha phat + omega shat ; r = s - omega t ; partial dots (rhat, r), (r, r)
__global__ void k_bicg_xr(int64_t n, const double* __restrict__ sc, const double* __restrict__ phat, const double* __restrict__ shat,
                          const double* __restrict__ s, const double* __restrict__ t, const double* __restrict__ rhat, double* __restrict__ x,
                          double* __restrict__ r, double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = sc[S_ALPHA], omega = sc[S_OMEGA];
    double d0 = 0.0, d1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * phat[i] + omega * shat[i];
        const double ri = s[i] - omega * t[i];
        r[i] = ri;
        d0 += rhat[i] * ri;
        d1 += ri * ri;
    }
    const double s0 = block_sum(d0, red);
    const double s1 = block_sum(d1, red);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s0;
        part[gridDim.x + blockIdx.x] = s1;
    }
}
// partial dots (a,b), (b,b)
__global__ void k_dot2(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ part) {
    __shared__ double red[32];
    double d0 = 0.0, d1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        d0 += a[i] * b[i];
        d1 += b[i] * b[i];
    }
    const double s0 = block_sum(d0, red);
    const double s1 = block_sum(d1, red);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s0;
        part[gridDim.x + blockIdx.x] = s1;
    }
}
// CG: x += alpha p ; r -= alpha q ; z = M^-1 r (if fused) ; partial dots (z,r), (r,r)
__global__ void k_cg_xr(int64_t n, const double* __restrict__ sc, const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
                        double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ z, int fused, double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = sc[S_ALPHA];
    double d0 = 0.0, d1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        if (fused) {
            const double zi = dinv ? ri * dinv[i] : ri;
            z[i] = zi;
            d0 += zi * ri;
            d1 += ri * ri;
        }
    }
    if (fused) {
        const double s0 = block_sum(d0, red);
        const double s1 = block_sum(d1, red);
        if (threadIdx.x == 0) {
            part[blockIdx.x] = s0;
            part[gridDim.x + blockIdx.x] = s1;
        }
    }
}
// CG: p = z + beta p
__global__ void k_cg_p(int64_t n, const double* __restrict__ sc, const double* __restrict__ z, double* __restrict__ p) {
    const double beta = sc[S_BETA];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}
// GMRES: partial dots (w, V_i) for i = 0..k-1 plus (w,w) in one pass; V is k vectors of length n, stride ld
__global__ void k_multidot(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ w, double* __restrict__ part) {
    __shared__ double red[32];
    for (int i0 = 0; i0 <= k; i0 += 4) {
        double d[4] = {0, 0, 0, 0};
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
            const double wj = w[j];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = i0 + q;
                if (i < k) d[q] += wj * V[(int64_t)i * ld + j];
                else if (i == k) d[q] += wj * wj;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double sres = block_sum(d[q], red);
            if (threadIdx.x == 0 && i0 + q <= k) part[(int64_t)(i0 + q) * gridDim.x + blockIdx.x] = sres;
            __syncthreads();
        }
    }
}
// GMRES: w -= sum_i hcoef[i] V_i
__global__ void k_multiaxpy(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ hcoef, double* __restrict__ w) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double wj = w[j];
        for (int i = 0; i < k; i++) wj -= hcoef[i] * V[(int64_t)i * ld + j];
        w[j] = wj;
    }
}
// y = a * x (+ y if acc)
__global__ void k_scale(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y, int acc) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) y[j] = acc ? y[j] + alpha * x[j] : alpha * x[j];
}
// z = dinv .* r (or copy if dinv is null)
__global__ void k_pmul(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) z[j] = dinv ? r[j] * dinv[j] : r[j];
}
// reciprocal point diagonal for the fused Jacobi preconditioner
template <int NS>
__global__ void k_dinv(int64_t Nown, const double* __restrict__ diagval, double* __restrict__ dinv, const SpmvArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Nown) return;
#pragma unroll
    for (int i = 0; i < NS; i++) dinv[r * NS + i] = 1.0 / diagval[(int64_t)a.idxD[i * NS + i] * Nown + r];
}

// ---- preconditioners ----------------------------------------------------------------------------------------
// point Jacobi: out = in / a_ii
template <int NS>
__global__ void k_jacobi(int64_t Nown, const double* __restrict__ diagval, const double* __restrict__ in, double* __restrict__ out, const SpmvArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Nown) return;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double d = diagval[(int64_t)a.idxD[i * NS + i] * Nown + r];
        out[r * NS + i] = in[r * NS + i] / d;
    }
}
// node-block Jacobi: factor (Doolittle LU without pivoting of the NS x NS node block, stored as its explicit inverse)
template <int NS>
__global__ void k_blockjacobi_setup(int64_t Nown, const double* __restrict__ diagval, double* __restrict__ inv, int32_t* __restrict__ flags, const SpmvArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Nown) return;
    double A[NS][NS], B[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; i++)
#pragma unroll
        for (int j = 0; j < NS; j++) {
            const int p = a.idxD[i * NS + j];
            A[i][j] = p >= 0 ? diagval[(int64_t)p * Nown + r] : 0.0;
            B[i][j] = (i == j) ? 1.0 : 0.0;
        }
    // Gauss-Jordan without pivoting (node blocks of FV Jacobians are diagonally dominant / M-matrix like)
#pragma unroll
    for (int c = 0; c < NS; c++) {
        const double piv = A[c][c];
        if (!(fabs(piv) > 0.0)) atomicOr(flags, 4);
        const double ip = 1.0 / piv;
#pragma unroll
        for (int j = 0; j < NS; j++) {
            A[c][j] *= ip;
            B[c][j] *= ip;
        }
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (i == c) continue;
            const double f = A[i][c];
#pragma unroll
            for (int j = 0; j < NS; j++) {
                A[i][j] -= f * A[c][j];
                B[i][j] -= f * B[c][j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++)
#pragma unroll
        for (int j = 0; j < NS; j++) inv[(int64_t)(i * NS + j) * Nown + r] = B[i][j];
}
template <int NS>
__global__ void k_blockjacobi_apply(int64_t Nown, const double* __restrict__ inv, const double* __restrict__ in, double* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Nown) return;
    double x[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) x[j] = in[r * NS + j];
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++) s += inv[(int64_t)(i * NS + j) * Nown + r] * x[j];
        out[r * NS + i] = s;
    }
}

// ---- Newton update + norms (K11) ---------------------------------------------------------------------------
// u -= damp * delta ; partials: max|delta|, sum|u_new|
__global__ void k_newton_update(int64_t n, double damp, const double* __restrict__ delta, double* __restrict__ u, double* __restrict__ part) {
    __shared__ double red[32];
    double mx = 0.0, s1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = delta[i];
        const double un = u[i] - damp * d;
        u[i] = un;
        mx = fmax(mx, fabs(d));
        s1 += fabs(un);
    }
    const double m = block_max(mx, red);
    const double s = block_sum(s1, red);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = m;
        part[gridDim.x + blockIdx.x] = s;
    }
}
__global__ void k_norms(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ part) {
    __shared__ double red[32];
    double mx = 0.0, s1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = b ? a[i] - b[i] : a[i];
        mx = fmax(mx, fabs(d));
        s1 += fabs(d);
    }
    const double m = block_max(mx, red);
    const double s = block_sum(s1, red);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = m;
        part[gridDim.x + blockIdx.x] = s;
    }
}

#define NS_DISPATCH(n, ...)                                   \
    switch (n) {                                              \
        case 1: { constexpr int NS = 1; __VA_ARGS__; } break;        \
        case 2: { constexpr int NS = 2; __VA_ARGS__; } break;        \
        case 3: { constexpr int NS = 3; __VA_ARGS__; } break;        \
        case 4: { constexpr int NS = 4; __VA_ARGS__; } break;        \
        case 5: { constexpr int NS = 5; __VA_ARGS__; } break;        \
        case 10: { constexpr int NS = 10; __VA_ARGS__; } break;      \
        default: throw std::string("number of species without device instantiation (supported: 1,2,3,4,5,10)"); \
    }

const int VEC_GRID = 148 * 8;

SpmvArgs make_spmv_args(vfvm_handle* h) {
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.sell_ptr = h->sell_ptr.p;
    a.colidx = h->colidx.p;
    a.offval = h->offval.p;
    a.diagval = h->diagval.p;
    a.nnz_sell = h->nnz_sell;
    a.Nown = h->Nown;
    a.nslices = h->ngroups;
    for (int b = 0; b < 100; b++) {
        a.idxF[b] = (signed char)(b < h->n * h->n ? h->idxF[b] : -1);
        a.idxD[b] = (signed char)(b < h->n * h->n ? h->idxD[b] : -1);
    }
    return a;
}

// reduce `nvals` x `nparts` block partials into sc[S_TMP0..] and apply the scalar recurrence `op`; with several ranks the
// all-reduce sits between the two
void finalize(vfvm_handle* h, const double* part, int nparts, int nvals, int op) {
    if (h->nranks <= 1) {
        k_finalize_op<<<1, 1024, 0, h->stream>>>(part, nparts, nvals, h->red.p, op, h->flags.p);
        h->launches++;
    } else if (h->peer_ok && nvals <= VFVM_PEER_RED_W) {
        k_finalize_op_peer<<<1, 1024, 0, h->stream>>>(part, nparts, nvals, h->red.p, op, h->flags.p, vfvm_peer_args_reduce(h));
        h->launches++;
    } else {
        k_finalize_op<<<1, 1024, 0, h->stream>>>(part, nparts, nvals, h->red.p, OP_NONE, h->flags.p);
        h->launches++;
        vfvm_comm_allreduce_sum(h, h->red.p + S_TMP0, nvals);
        if (op != OP_NONE) {
            k_scalar<<<1, 1, 0, h->stream>>>(op, h->red.p, h->flags.p);
            h->launches++;
        }
    }
}

// which matrices take the bulk-copy kernel: VFVM_SPMV_BULK = 0 (never), 1 (default: coupled species, the finest level only), 2 (every
// species count, the finest level only)
static int spmv_bulk_mode() {
    static const int mode = getenv("VFVM_SPMV_BULK") ? atoi(getenv("VFVM_SPMV_BULK")) : 1;
    return mode;
}

template <int NS, bool DIAGMASK, bool PEER, class VT>
bool launch_spmv_bulk(vfvm_handle* h, SpmvArgs& a, int op, const LevelHalo* lh) {
    const int mode = spmv_bulk_mode();
    if (mode == 0 || a.nslices < 4096) return false;  // small levels are latency bound: nothing to stream
    if (mode == 1 && NS == 1) return false;
    const int cF = DIAGMASK ? NS : h->cF;
    const size_t smem = (size_t)(BULK_THREADS / 32) * BULK_STAGES * (BULK_CH * 32 * 4 + (size_t)cF * BULK_CH * 32 * sizeof(VT));
    if (smem > 100 * 1024) return false;  // many planes: fewer than two blocks per SM would fit, the register-staged kernel is the better one
    auto kern = k_spmv_bulk<NS, DIAGMASK, PEER, VT>;
    static int occ = 0;
    static size_t occ_smem = 0;
    if (occ == 0 || occ_smem != smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BULK_THREADS, smem));
        occ_smem = smem;
        if (occ < 1) {
            occ = 0;
            return false;
        }
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::max(1, std::min(cdiv(a.nslices, BULK_THREADS / 32), nsm * occ));
    if (a.w) {
        if (h->work[10].n < (size_t)2 * grid) h->work[10].alloc((size_t)2 * grid);
        a.part = h->work[10].p;
    }
    PeerArgs P;
    if constexpr (PEER) P = lh ? vfvm_peer_args_halo_level(h, *lh) : vfvm_peer_args_halo(h);
    else memset(&P, 0, sizeof(P));
    kern<<<grid, BULK_THREADS, smem, h->stream>>>(a, P, cF);
    h->launches++;
    if (a.w) finalize(h, a.part, grid, 2, op);
    return true;
}

template <int NS, bool DIAGMASK, bool PEER, class VT>
void launch_spmv_kv(vfvm_handle* h, SpmvArgs& a, int op, const LevelHalo* lh);

template <int NS, bool DIAGMASK, bool PEER>
void launch_spmv_k(vfvm_handle* h, SpmvArgs& a, int op, const LevelHalo* lh = nullptr) {
    if (a.offval32) launch_spmv_kv<NS, DIAGMASK, PEER, float>(h, a, op, lh);
    else launch_spmv_kv<NS, DIAGMASK, PEER, double>(h, a, op, lh);
}

template <int NS, bool DIAGMASK, bool PEER, class VT>
void launch_spmv_kv(vfvm_handle* h, SpmvArgs& a, int op, const LevelHalo* lh) {
    if (launch_spmv_bulk<NS, DIAGMASK, PEER, VT>(h, a, op, lh)) return;
    auto kern = k_spmv<NS, DIAGMASK, PEER, VT>;
    static int occ = 0;
    if (occ == 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LS_THREADS, 0));
        if (occ < 1) throw std::string("SpMV kernel cannot be launched");
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::max(1, std::min(cdiv(a.nslices, LS_THREADS / 32), nsm * occ));
    if (a.w) {
        if (h->work[10].n < (size_t)2 * grid) h->work[10].alloc((size_t)2 * grid);
        a.part = h->work[10].p;
    }
    PeerArgs P;
    if constexpr (PEER) P = lh ? vfvm_peer_args_halo_level(h, *lh) : vfvm_peer_args_halo(h);  // the grid is one resident wave, so the block that finishes the push last can raise the flags
    else memset(&P, 0, sizeof(P));
    kern<<<grid, LS_THREADS, 0, h->stream>>>(a, P);
    h->launches++;
    if (a.w) finalize(h, a.part, grid, 2, op);
}

template <int NS>
void launch_spmv(vfvm_handle* h, SpmvArgs& a, int op) {
    bool diagmask = (h->cF == NS && h->cD == NS);
    for (int i = 0; i < NS && diagmask; i++) diagmask = (h->idxF[i * NS + i] == i && h->idxD[i * NS + i] == i);
    const bool peer = h->nranks > 1 && h->peer_ok && !h->nb_ranks.empty();
    if (diagmask) {
        if (peer) launch_spmv_k<NS, true, true>(h, a, op);
        else launch_spmv_k<NS, true, false>(h, a, op);
    } else {
        if (peer) launch_spmv_k<NS, false, true>(h, a, op);
        else launch_spmv_k<NS, false, false>(h, a, op);
    }
}

// y = A x (+ fused dots (y,w), (y,y) -> sc[S_TMP0], sc[S_TMP1], then scalar recurrence `op`)
void spmv(vfvm_handle* h, double* x, double* y, const double* w, int op = OP_NONE) {
    if (h->nranks > 1 && !h->peer_ok) vfvm_halo_exchange_ptr(h, x);  // NCCL transport; with peer mailboxes the exchange is part of the SpMV kernel
    SpmvArgs a = make_spmv_args(h);
    a.x = x;
    a.y = y;
    a.w = w;
    NS_DISPATCH(h->n, launch_spmv<NS>(h, a, op));
}

void precond_setup(vfvm_handle* h) {
    const int64_t Nown = h->Nown;
    if (h->precon == VFVM_PRECON_JACOBI) {
        h->pc_diag.alloc((size_t)h->n * Nown);
        SpmvArgs a = make_spmv_args(h);
        NS_DISPATCH(h->n, (k_dinv<NS><<<cdiv(Nown, 256), 256, 0, h->stream>>>(Nown, h->diagval.p, h->pc_diag.p, a)));
        h->launches++;
    } else if (h->precon == VFVM_PRECON_BLOCKJACOBI) {
        h->pc_diag.alloc((size_t)h->n * h->n * Nown);
        SpmvArgs a = make_spmv_args(h);
        NS_DISPATCH(h->n, (k_blockjacobi_setup<NS><<<cdiv(Nown, 128), 128, 0, h->stream>>>(Nown, h->diagval.p, h->pc_diag.p, h->flags.p, a)));
        h->launches++;
    } else if (h->precon == VFVM_PRECON_ILU0 || h->precon == VFVM_PRECON_ILU0_MC) {
        vfvm_ilu0_setup(h);
    } else if (h->precon == VFVM_PRECON_AMG) {
        vfvm_amg_setup(h);
    }
    h->precon_valid = true;
}

// preconditioners that are not fused into the vector kernels
void precond_apply(vfvm_handle* h, const double* in, double* out) {
    const int64_t Nown = h->Nown;
    const int64_t nd = Nown * h->n;
    switch (h->precon) {
        case VFVM_PRECON_NONE: CK(cudaMemcpyAsync(out, in, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream)); break;
        case VFVM_PRECON_JACOBI:
            k_pmul<<<VEC_GRID, LS_THREADS, 0, h->stream>>>(nd, h->pc_diag.p, in, out);
            h->launches++;
            break;
        case VFVM_PRECON_BLOCKJACOBI:
            NS_DISPATCH(h->n, (k_blockjacobi_apply<NS><<<cdiv(Nown, 128), 128, 0, h->stream>>>(Nown, h->pc_diag.p, in, out)));
            h->launches++;
            break;
        case VFVM_PRECON_ILU0:
        case VFVM_PRECON_ILU0_MC: vfvm_ilu0_apply(h, in, out); break;
        case VFVM_PRECON_AMG: vfvm_amg_apply(h, in, out); break;
    }
}

// stopping test without stalling the stream: the residual norm of iteration k is copied to pinned memory behind an event
// and inspected after iteration k+1 has been enqueued (at most one extra iteration is performed)
struct AsyncNorm {
    vfvm_handle* h;
    cudaEvent_t ev[2];
    int pending[2] = {-1, -1};
    explicit AsyncNorm(vfvm_handle* hh) : h(hh) {
        ev[0] = hh->ev3;
        ev[1] = hh->ev4;
    }
    void post(int it, int slot_rr) {
        const int s = it & 1;
        CK(cudaMemcpyAsync(h->red_host + 2 * s, h->red.p + slot_rr, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->flags_host + s, h->flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(ev[s], h->stream));
        pending[s] = it;
    }
    // returns true and fills rr / flags if iteration `it` has been posted
    bool fetch(int it, double& rr, int& flags) {
        const int s = it & 1;
        if (it < 0 || pending[s] != it) return false;
        CK(cudaEventSynchronize(ev[s]));
        rr = h->red_host[2 * s];
        flags = h->flags_host[s];
        pending[s] = -1;
        return true;
    }
};


// ---- one Krylov iteration as a CUDA graph ---------------------------------------------------------------------------------------
// Every kernel of an iteration (SpMV with its fused halo push / dots, the vector updates, the whole AMG cycle, the single-block
// reductions with their peer all-reduce) takes arguments that do not change from one iteration to the next: the Krylov scalars and
// the peer sequence counters live in device memory.  The first two iterations of a solve run eagerly (allocations, occupancy
// queries), the third is captured once per (method, preconditioner, buffers) signature and kept in the handle; every later
// iteration -- of this solve and of the following Newton steps -- is one cudaGraphLaunch.  With several ranks this removes the
// launch gaps of the ~50 small coarse-level kernels of a distributed AMG cycle, which bounded the 8-GPU step in round 1.
uint64_t iteration_signature(vfvm_handle* h, const double* b, const double* x, const double* part) {
    uint64_t sig = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((uint64_t)h->krylov);
    mix((uint64_t)h->precon);
    mix((uint64_t)h->graph_epoch);
    mix((uint64_t)h->Nown);
    mix((uint64_t)h->nnz_sell);
    mix((uint64_t)(uintptr_t)b);
    mix((uint64_t)(uintptr_t)x);
    mix((uint64_t)(uintptr_t)part);
    mix((uint64_t)(uintptr_t)h->offval.p);
    mix((uint64_t)(uintptr_t)h->diagval.p);
    mix((uint64_t)(uintptr_t)h->pc_diag.p);
    mix((uint64_t)(h->peer_ok ? 1 : 0));
    for (int i = 0; i < 8; i++) mix((uint64_t)(uintptr_t)h->work[i].p);
    mix((uint64_t)(uintptr_t)h->work[10].p);
    return sig;
}

struct IterationGraph {
    vfvm_handle* h;
    const double *b, *x, *part;
    uint64_t sig = 0;
    bool enabled;
    IterationGraph(vfvm_handle* hh, const double* b_, const double* x_, const double* part_) : h(hh), b(b_), x(x_), part(part_) {
        static const bool off = getenv("VFVM_NO_KRYLOV_GRAPH") != nullptr;
        // ILU applications launch one kernel per dependency level from host-side level lists: eager; NCCL transport: eager
        enabled = !off && (h->nranks <= 1 || h->peer_ok) && h->precon != VFVM_PRECON_ILU0 && h->precon != VFVM_PRECON_ILU0_MC &&
                  !(h->precon == VFVM_PRECON_AMG && h->amg_nccl_in_cycle);
    }
    template <class Body>
    void run(int it, Body& body) {
        if (!enabled || it <= 2) {
            body();
            return;
        }
        if (!sig) sig = iteration_signature(h, b, x, part);  // after the eager iterations: every buffer exists now
        if (h->iter_graph && h->iter_graph_sig == sig) {
            CK(cudaGraphLaunch((cudaGraphExec_t)h->iter_graph, h->stream));
            h->launches += h->iter_graph_launches;
            return;
        }
        if (h->iter_graph) {
            cudaGraphExecDestroy((cudaGraphExec_t)h->iter_graph);
            h->iter_graph = nullptr;
        }
        const int64_t before = h->launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->in_capture = true;
        try {
            body();
        } catch (...) {
            h->in_capture = false;
            cudaStreamEndCapture(h->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        h->in_capture = false;
        CK(cudaStreamEndCapture(h->stream, &graph));
        h->iter_graph_launches = h->launches - before;
        cudaGraphExec_t exec = nullptr;
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        h->iter_graph = exec;
        h->iter_graph_sig = sig;
        CK(cudaGraphLaunch(exec, h->stream));
    }
};

}  // namespace

void vfvm_spmv_impl(vfvm_handle* h, const double* x, double* y, bool planes32) {
    if (!planes32 || !h->offval32.p) {
        spmv(h, const_cast<double*>(x), y, nullptr);
        return;
    }
    if (h->nranks > 1 && !h->peer_ok) vfvm_halo_exchange_ptr(h, const_cast<double*>(x));
    SpmvArgs a = make_spmv_args(h);
    a.x = x;
    a.y = y;
    a.w = nullptr;
    a.offval32 = h->offval32.p;
    NS_DISPATCH(h->n, launch_spmv<NS>(h, a, OP_NONE));
}

// y = A x for any matrix in the DBSR / SELL-32 layout described by `a` (AMG levels): rank-local, no halo exchange, no dots
void vfvm_spmv_level(vfvm_handle* h, SpmvArgs a, const double* x, double* y) {
    a.x = x;
    a.y = y;
    a.w = nullptr;
    bool diagmask = (h->cF == h->n && h->cD == h->n);
    for (int i = 0; i < h->n && diagmask; i++) diagmask = (h->idxF[i * h->n + i] == i && h->idxD[i * h->n + i] == i);
    if (diagmask) {
        NS_DISPATCH(h->n, (launch_spmv_k<NS, true, false>(h, a, OP_NONE)));
    } else {
        NS_DISPATCH(h->n, (launch_spmv_k<NS, false, false>(h, a, OP_NONE)));
    }
}
// the same with the halo of x refreshed first (coarser AMG levels of a distributed hierarchy): over the peer mailboxes the exchange
// is part of the SpMV kernel, otherwise an NCCL exchange precedes it
void vfvm_spmv_level_halo(vfvm_handle* h, SpmvArgs a, LevelHalo& lh, double* x, double* y) {
    if (!h->peer_ok) {
        vfvm_halo_exchange_level(h, lh, x);
        vfvm_spmv_level(h, a, x, y);
        return;
    }
    a.x = x;
    a.y = y;
    a.w = nullptr;
    bool diagmask = (h->cF == h->n && h->cD == h->n);
    for (int i = 0; i < h->n && diagmask; i++) diagmask = (h->idxF[i * h->n + i] == i && h->idxD[i * h->n + i] == i);
    if (diagmask) {
        NS_DISPATCH(h->n, (launch_spmv_k<NS, true, true>(h, a, OP_NONE, &lh)));
    } else {
        NS_DISPATCH(h->n, (launch_spmv_k<NS, false, true>(h, a, OP_NONE, &lh)));
    }
}
SpmvArgs vfvm_spmv_args(vfvm_handle* h) { return make_spmv_args(h); }
// explicit inverses of the n x n diagonal blocks of a level (planes i*n+j of N entries)
void vfvm_blockinv_level(vfvm_handle* h, const SpmvArgs& a, int64_t N, const double* diagval, double* inv) {
    NS_DISPATCH(h->n, (k_blockjacobi_setup<NS><<<cdiv(N, 128), 128, 0, h->stream>>>(N, diagval, inv, h->flags.p, a)));
    h->launches++;
}

extern "C" int vfvm_linsolve_setup(vfvm_handle* h, int krylov, int precon, int gmres_restart) {
    if (!h) return VFVM_ERR_ARG;
    if (krylov < VFVM_KRYLOV_BICGSTAB || krylov > VFVM_KRYLOV_GMRES) return vfvm_fail(h, VFVM_ERR_ARG, "unknown Krylov method");
    if (precon < VFVM_PRECON_NONE || precon > VFVM_PRECON_AMG) return vfvm_fail(h, VFVM_ERR_ARG, "unknown preconditioner");
    h->krylov = krylov;
    h->precon = precon;
    h->gmres_restart = gmres_restart > 0 ? std::min(gmres_restart, 100) : 30;
    h->precon_valid = false;
    h->graph_epoch++;
    return VFVM_OK;
}

// solves A * UPDATE = RESIDUAL, initial guess 0 (LinearSolve semantics: u .= sol.u, src/vfvm_linsolve.jl:46-48)
extern "C" int vfvm_linsolve(vfvm_handle* h, double abstol, double reltol, int maxiters, int reuse_precs, int* iters, double* resnorm) {
    if (!h || !h->have_pattern) return vfvm_fail(h, VFVM_ERR_STATE, "vfvm_build_pattern has not been called");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        cudaStream_t st = h->stream;
        const int64_t nd = h->Nown * h->n, nall = h->N * h->n;
        double* b = h->vec[VFVM_VEC_RESIDUAL].p;
        double* x = h->vec[VFVM_VEC_UPDATE].p;
        for (int i = 0; i < 8; i++)
            if (h->work[i].n != (size_t)nall) {
                h->work[i].alloc(nall);
                CK(cudaMemsetAsync(h->work[i].p, 0, nall * sizeof(double), st));
            }
        if (h->work[11].n < (size_t)128 * VEC_GRID) h->work[11].alloc((size_t)128 * VEC_GRID);
        double* part = h->work[11].p;
        CK(cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t), st));

        CK(cudaEventRecord(h->ev0, st));
        // reuse_precs (factorize_every_newtonstep = false, src/vfvm_solver.jl:99) keeps an ILU factorisation across Newton steps, as the
        // reference keeps its preconditioner.  Jacobi, node-block Jacobi and the AMG numeric phase cost less than one Krylov iteration on
        // the device and the AMG cycle multiplies with the CURRENT finest-level Jacobian, so they are always refreshed (a stale Galerkin
        // hierarchy under a new finest level is not a consistent preconditioner: measured stagnation on the bipolar system).
        const bool keep = reuse_precs && h->precon_valid && (h->precon == VFVM_PRECON_ILU0 || h->precon == VFVM_PRECON_ILU0_MC);
        if (!keep) precond_setup(h);
        CK(cudaEventRecord(h->ev1, st));

        const bool fusedpc = (h->precon == VFVM_PRECON_NONE || h->precon == VFVM_PRECON_JACOBI);
        const double* dinv = h->precon == VFVM_PRECON_JACOBI ? h->pc_diag.p : nullptr;
        CK(cudaMemsetAsync(x, 0, nall * sizeof(double), st));
        int it = 0, done_it = 0;
        double rr = 0.0, bb = 0.0, tol = 0.0;
        bool converged = false, breakdown = false;
        AsyncNorm an(h);
        auto check = [&](int k) {  // inspects iteration k if posted
            double v;
            int fl;
            if (!an.fetch(k, v, fl)) return;
            rr = v;
            done_it = k;
            if (fl & 256) throw std::string("peer exchange timed out: a neighbouring rank did not deliver its halo / partial sums");
            if (v != v || (fl & 6)) breakdown = true;
            else if (sqrt(v) <= tol) converged = true;
        };
        if (h->krylov == VFVM_KRYLOV_BICGSTAB) {
            double *r = h->work[0].p, *rhat = h->work[1].p, *p = h->work[2].p, *v = h->work[3].p, *s = h->work[4].p, *t = h->work[5].p, *phat = h->work[6].p,
                   *shat = h->work[7].p;
            CK(cudaMemcpyAsync(r, b, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(rhat, b, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemsetAsync(p, 0, nd * sizeof(double), st));
            CK(cudaMemsetAsync(v, 0, nd * sizeof(double), st));
            k_dot2<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, r, r, part);
            h->launches++;
            finalize(h, part, VEC_GRID, 2, OP_BICG_INIT);
            an.post(0, S_BB);
            check(0);
            bb = rr;
            tol = fmax(abstol, reltol * sqrt(bb));
            converged = sqrt(rr) <= tol;
            auto body = [&]() {
                k_bicg_p<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, h->red.p, r, v, p, dinv, fusedpc ? phat : nullptr, rhat);
                h->launches++;
                if (!fusedpc) precond_apply(h, p, phat);
                spmv(h, phat, v, rhat, OP_BICG_ALPHA);  // (v, rhat)
                k_bicg_s<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, h->red.p, r, v, s, dinv, fusedpc ? shat : nullptr);
                h->launches++;
                if (!fusedpc) precond_apply(h, s, shat);
                spmv(h, shat, t, s, OP_BICG_OMEGA);  // (t, s), (t, t)
                k_bicg_xr<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, h->red.p, phat, shat, s, t, rhat, x, r, part);
                h->launches++;
                finalize(h, part, VEC_GRID, 2, OP_BICG_NEXT);  // (rhat, r), (r, r)
            };
            IterationGraph ig(h, b, x, part);
            while (!converged && !breakdown && it < maxiters) {
                it++;
                ig.run(it, body);
                an.post(it, S_RR);
                check(it - 1);
            }
            if (!converged && !breakdown) check(it);
        } else if (h->krylov == VFVM_KRYLOV_CG) {
            double *r = h->work[0].p, *z = h->work[1].p, *p = h->work[2].p, *q = h->work[3].p;
            CK(cudaMemcpyAsync(r, b, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
            if (fusedpc) {
                k_pmul<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, dinv, r, z);
                h->launches++;
            } else {
                precond_apply(h, r, z);
            }
            CK(cudaMemcpyAsync(p, z, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
            k_dot2<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, z, r, part);  // (z,r), (r,r)
            h->launches++;
            finalize(h, part, VEC_GRID, 2, OP_CG_INIT);
            an.post(0, S_BB);
            check(0);
            bb = rr;
            tol = fmax(abstol, reltol * sqrt(bb));
            converged = sqrt(rr) <= tol;
            auto body = [&]() {
                spmv(h, p, q, p, OP_CG_ALPHA);  // (q, p)
                k_cg_xr<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, h->red.p, p, q, x, r, dinv, z, fusedpc ? 1 : 0, part);
                h->launches++;
                if (!fusedpc) {
                    precond_apply(h, r, z);
                    k_dot2<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, z, r, part);
                    h->launches++;
                }
                finalize(h, part, VEC_GRID, 2, OP_CG_NEXT);
                k_cg_p<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, h->red.p, z, p);
                h->launches++;
            };
            IterationGraph ig(h, b, x, part);
            while (!converged && !breakdown && it < maxiters) {
                it++;
                ig.run(it, body);
                an.post(it, S_RR);
                check(it - 1);
            }
            if (!converged && !breakdown) check(it);
        } else {
            // restarted GMRES(m), right preconditioned, classical Gram-Schmidt with one fused multi-dot pass (+ one
            // re-orthogonalisation pass), Givens rotations on the host (m+1 scalars per iteration cross PCIe)
            const int m = h->gmres_restart;
            if (h->work[8].n != (size_t)(m + 1) * nall) h->work[8].alloc((size_t)(m + 1) * nall);
            if (h->work[9].n < (size_t)(m + 2)) h->work[9].alloc((size_t)(m + 2));
            double* V = h->work[8].p;
            double* hdev = h->work[9].p;
            double *w = h->work[0].p, *zv = h->work[1].p, *r = h->work[2].p;
            std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m, 0.0), sn(m, 0.0), gvec(m + 1, 0.0), hcol(m + 2, 0.0), y(m, 0.0);
            auto dots_to_host = [&](int k, const double* wv) {  // hcol[0..k-1] = (w, V_i), hcol[k] = (w,w)
                k_multidot<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, k, V, nall, wv, part);
                k_finalize<<<1, 1024, 0, st>>>(part, VEC_GRID, k + 1, hdev);
                h->launches += 2;
                vfvm_comm_allreduce_sum(h, hdev, k + 1);
                CK(cudaMemcpyAsync(hcol.data(), hdev, (k + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
            };
            CK(cudaMemcpyAsync(r, b, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
            dots_to_host(0, r);
            bb = hcol[0];
            rr = bb;
            tol = fmax(abstol, reltol * sqrt(bb));
            converged = sqrt(rr) <= tol;
            while (!converged && !breakdown && it < maxiters) {
                const double beta = sqrt(rr);
                k_scale<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, 1.0 / beta, r, V, 0);
                h->launches++;
                std::fill(gvec.begin(), gvec.end(), 0.0);
                gvec[0] = beta;
                int k = 0;
                for (; k < m && it < maxiters && !converged; k++) {
                    it++;
                    double* vk = V + (int64_t)k * nall;
                    if (fusedpc) {
                        k_pmul<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, dinv, vk, zv);
                        h->launches++;
                    } else {
                        precond_apply(h, vk, zv);
                    }
                    spmv(h, zv, w, nullptr);
                    for (int pass = 0; pass < 2; pass++) {  // CGS2
                        dots_to_host(k + 1, w);
                        CK(cudaMemcpyAsync(hdev, hcol.data(), (k + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
                        k_multiaxpy<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, k + 1, V, nall, hdev, w);
                        h->launches++;
                        for (int i = 0; i <= k; i++) H[(size_t)i * m + k] += hcol[i];
                    }
                    dots_to_host(0, w);
                    const double hk1 = sqrt(hcol[0]);
                    if (hk1 != hk1) {
                        breakdown = true;
                        break;
                    }
                    H[(size_t)(k + 1) * m + k] = hk1;
                    if (hk1 > 0.0) {
                        k_scale<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, 1.0 / hk1, w, V + (int64_t)(k + 1) * nall, 0);
                        h->launches++;
                    }
                    for (int i = 0; i < k; i++) {  // previous rotations
                        const double t1 = cs[i] * H[(size_t)i * m + k] + sn[i] * H[(size_t)(i + 1) * m + k];
                        H[(size_t)(i + 1) * m + k] = -sn[i] * H[(size_t)i * m + k] + cs[i] * H[(size_t)(i + 1) * m + k];
                        H[(size_t)i * m + k] = t1;
                    }
                    const double a1 = H[(size_t)k * m + k], a2 = H[(size_t)(k + 1) * m + k], den = hypot(a1, a2);
                    cs[k] = den > 0 ? a1 / den : 1.0;
                    sn[k] = den > 0 ? a2 / den : 0.0;
                    H[(size_t)k * m + k] = den;
                    H[(size_t)(k + 1) * m + k] = 0.0;
                    gvec[k + 1] = -sn[k] * gvec[k];
                    gvec[k] = cs[k] * gvec[k];
                    rr = gvec[k + 1] * gvec[k + 1];
                    converged = fabs(gvec[k + 1]) <= tol;
                }
                // x += M^-1 V y with H y = g
                for (int i = k - 1; i >= 0; i--) {
                    double sum = gvec[i];
                    for (int j = i + 1; j < k; j++) sum -= H[(size_t)i * m + j] * y[j];
                    y[i] = sum / H[(size_t)i * m + i];
                }
                CK(cudaMemsetAsync(w, 0, nd * sizeof(double), st));
                for (int i = 0; i < k; i++) {
                    k_scale<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, y[i], V + (int64_t)i * nall, w, 1);
                    h->launches++;
                }
                if (fusedpc) {
                    k_pmul<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, dinv, w, zv);
                    h->launches++;
                } else {
                    precond_apply(h, w, zv);
                }
                k_scale<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, 1.0, zv, x, 1);
                h->launches++;
                std::fill(H.begin(), H.end(), 0.0);
                if (!converged && !breakdown && it < maxiters) {  // true residual for the restart
                    spmv(h, x, w, nullptr);
                    CK(cudaMemcpyAsync(r, b, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
                    k_scale<<<VEC_GRID, LS_THREADS, 0, st>>>(nd, -1.0, w, r, 1);
                    h->launches++;
                    dots_to_host(0, r);
                    rr = hcol[0];
                    converged = sqrt(rr) <= tol;
                }
            }
        }
        CK(cudaEventRecord(h->ev2, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->times[VFVM_TIME_LINSOLVE_SETUP] = ms;
        CK(cudaEventElapsedTime(&ms, h->ev1, h->ev2));
        h->times[VFVM_TIME_LINSOLVE_SOLVE] = ms;
        *iters = it;
        *resnorm = sqrt(rr);
        h->last_converged = converged ? 1 : 0;
        h->last_rhsnorm = sqrt(bb);
        if (breakdown) {
            h->err = "Krylov breakdown (zero pivot / NaN)";
            return VFVM_ERR_LINSOLVE;
        }
        // like Krylov.jl under LinearSolve, hitting maxiters is not an error: the Newton loop judges the update
    })
    return VFVM_OK;
}

// outcome of the last vfvm_linsolve: hitting maxiters is not an error of the call (Krylov.jl under LinearSolve behaves the same and the
// Newton loop judges the update), but a caller that stands in for a DIRECT solve must know whether the tolerance was reached
extern "C" int vfvm_linsolve_status(vfvm_handle* h, int* converged, double* rhs_norm) {
    if (!h) return VFVM_ERR_ARG;
    if (converged) *converged = h->last_converged;
    if (rhs_norm) *rhs_norm = h->last_rhsnorm;
    return VFVM_OK;
}

extern "C" int vfvm_spmv(vfvm_handle* h, const double* x, double* y, int memspace) {
    if (!h || !h->have_pattern) return vfvm_fail(h, VFVM_ERR_STATE, "vfvm_build_pattern has not been called");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        const int64_t nd = h->Nown * h->n, nall = h->N * h->n;
        if (memspace == VFVM_DEVICE) {
            spmv(h, const_cast<double*>(x), y, nullptr);
        } else {
            DevBuf<double> dx, dy;
            dx.upload(x, nall, h->stream);
            dy.alloc(nd);
            spmv(h, dx.p, dy.p, nullptr);
            CK(cudaMemcpyAsync(y, dy.p, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        return vfvm_peer_check(h);
    })
}

static int norms_impl(vfvm_handle* h, const double* a, const double* b, double* norm_inf, double* norm1) {
    CK(cudaSetDevice(h->device));
    const int64_t nd = h->Nown * h->n;
    if (h->work[11].n < (size_t)4 * VEC_GRID) h->work[11].alloc((size_t)4 * VEC_GRID);
    double* part = h->work[11].p;
    k_norms<<<VEC_GRID, LS_THREADS, 0, h->stream>>>(nd, a, b, part);
    k_finalize_max<<<1, 1024, 0, h->stream>>>(part, VEC_GRID, h->red.p + S_TMP0);
    k_finalize<<<1, 1024, 0, h->stream>>>(part + VEC_GRID, VEC_GRID, 1, h->red.p + S_TMP1);
    h->launches += 3;
    vfvm_comm_allreduce_max(h, h->red.p + S_TMP0, 1);
    vfvm_comm_allreduce_sum(h, h->red.p + S_TMP1, 1);
    CK(cudaMemcpyAsync(h->red_host, h->red.p + S_TMP0, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (norm_inf) *norm_inf = h->red_host[0];
    if (norm1) *norm1 = h->red_host[1];
    return vfvm_peer_check(h);
}

extern "C" int vfvm_newton_update(vfvm_handle* h, double damp, double* update_norm_inf, double* solution_norm1) {
    if (!h || !h->have_pattern) return vfvm_fail(h, VFVM_ERR_STATE, "vfvm_build_pattern has not been called");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        const int64_t nd = h->Nown * h->n;
        if (h->work[11].n < (size_t)4 * VEC_GRID) h->work[11].alloc((size_t)4 * VEC_GRID);
        double* part = h->work[11].p;
        vfvm_zero_inactive(h, h->vec[VFVM_VEC_UPDATE].p);
        k_newton_update<<<VEC_GRID, LS_THREADS, 0, h->stream>>>(nd, damp, h->vec[VFVM_VEC_UPDATE].p, h->vec[VFVM_VEC_SOLUTION].p, part);
        k_finalize_max<<<1, 1024, 0, h->stream>>>(part, VEC_GRID, h->red.p + S_TMP0);
        k_finalize<<<1, 1024, 0, h->stream>>>(part + VEC_GRID, VEC_GRID, 1, h->red.p + S_TMP1);
        h->launches += 3;
        vfvm_comm_allreduce_max(h, h->red.p + S_TMP0, 1);
        vfvm_comm_allreduce_sum(h, h->red.p + S_TMP1, 1);
        if (h->nranks > 1) vfvm_halo_exchange_ptr(h, h->vec[VFVM_VEC_SOLUTION].p);
        CK(cudaMemcpyAsync(h->red_host, h->red.p + S_TMP0, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        if (update_norm_inf) *update_norm_inf = h->red_host[0];
        if (solution_norm1) *solution_norm1 = h->red_host[1];
        return vfvm_peer_check(h);
    })
}

extern "C" int vfvm_vector_norms(vfvm_handle* h, int which, double* norm_inf, double* norm1) {
    if (!h || !h->have_pattern || which < 0 || which > 3) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id or no pattern");
    VFVM_TRY(h, { return norms_impl(h, h->vec[which].p, nullptr, norm_inf, norm1); })
}

extern "C" int vfvm_vector_diffnorm(vfvm_handle* h, int which_a, int which_b, double* norm_inf) {
    if (!h || !h->have_pattern || which_a < 0 || which_a > 3 || which_b < 0 || which_b > 3) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id or no pattern");
    VFVM_TRY(h, { return norms_impl(h, h->vec[which_a].p, h->vec[which_b].p, norm_inf, nullptr); })
}
