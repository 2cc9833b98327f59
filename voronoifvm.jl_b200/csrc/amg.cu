// Aggregation algebraic multigrid preconditioner (VFVM_PRECON_AMG).
//
// The reference reaches multigrid through `precs = AMGPreconBuilder()` / `SmoothedAggregationPreconBuilder()` of
// AMGCLWrap / AlgebraicMultigrid (examples/DevEx003_Solvers.jl:149-169, examples/DevEx004_EquationBlock3D.jl:225-251);
// this is the device twin: a V-cycle of plain (piecewise constant) aggregation on the NODE graph, species kept as n x n
// blocks, so that every level is again a matrix in the DBSR / SELL-32 layout and reuses the SpMV kernel.
//
//   hierarchy (once per sparsity pattern, independent of the matrix values):
//     strength graph = the Voronoi edge factors sigma/h (nzfac) -- symmetric and value independent; on coarse levels their
//     Galerkin sums.  Aggregates = root + strong neighbours, roots chosen in rounds as the maxima of a fixed hash priority
//     among the still eligible nodes within distance two (no two roots share a neighbour => conflict-free, deterministic);
//     left-over nodes join the neighbouring aggregate they are most strongly tied to.  Nodes whose whole diagonal block is a
//     Dirichlet penalty (1e30) stay out of the coarse problem.
//     coarse pattern: sort the fine entries by (aggregate of row, aggregate of column) -- one CUB radix sort -- and
//     run-length encode; the sorted entry list is kept as the gather map of the numeric Galerkin product.
//   numeric phase (every new Jacobian): A_c(I,J) = sum of the fine blocks (K,L), K in I, L in J, plane by plane, as a GATHER
//     in the fixed sorted order (no atomics => bitwise reproducible); block inverses of the diagonal for the smoother.
//   cycle: damped block-Jacobi pre-/post-smoothing (symmetric, so CG stays applicable), residual fused into the restriction,
//     over-weighted coarse correction (plain aggregation under-estimates smooth corrections), four sweeps on the coarsest level;
//     optionally levels 1..wdepth are visited twice (W-cycle on the top of the hierarchy); the two finest-level SpMVs read an fp32
//     copy of the off-diagonal planes.
// With several ranks the hierarchy is distributed: aggregates never cross a partition boundary (every rank aggregates its owned
// nodes), but the couplings across the boundary are kept on every level.  A rank learns the aggregate of each of its halo nodes
// from the owner (one halo exchange of the aggregate ids per level); the distinct aggregates per neighbour, in ascending order,
// are the halo of the next level -- and the owner derives the matching send list from the same ids without a second exchange.
// Restriction, prolongation, Galerkin product and smoothing are rank-local; every SpMV of the cycle refreshes the halo of its
// input (fused into the SpMV kernel over the peer mailboxes, or an NCCL exchange in front of it).  From the first level whose stored
// values stay below a threshold every rank holds the WHOLE level and the sub-hierarchy under it (Amg::repl_level, replicate_level):
// one all-gather of the restricted right-hand side per cycle replaces the exchanges of those levels.
// VFVM_AMG_LOCAL=1 drops the halo couplings instead (block-Jacobi across ranks with AMG inside).
#include <algorithm>
#include <cub/cub.cuh>

#include "vfvm_internal.h"

namespace {

#define FUSE_THREADS 512
#define FUSE_SMALL_N 512
enum { FOP_SMOOTH_FIRST = 1, FOP_SPMV, FOP_SMOOTH, FOP_RESTRICT, FOP_WRES, FOP_WADD, FOP_PROLONG };

struct LevelDev {
    int64_t N, Nvec, nnz_sell, Nc;
    int nslices;
    const int32_t *sell_ptr, *colidx;
    const double *offval, *diagval, *binv;
    const int32_t *agg, *agg_ptr, *agg_nodes;
    double *x, *t, *b, *b2, *x2;
    // several ranks: this level's exchange lists (neighbour slots as on level 0)
    int64_t send_ptr[VFVM_PEER_MAX + 1], recv_ptr[VFVM_PEER_MAX + 1];
    const int32_t* send_idx;
};
struct FuseOp {
    int code, lvl, bsel, small;  // bsel: 0 = the level's b, 1 = its b2; small: executed by block 0 alone
};
struct FuseArgs {
    const LevelDev* lev;
    const FuseOp* prog;
    int nops, distributed;
    double omega, alpha;
    unsigned long long *bar_ctr, *bar_gen;
    PeerArgs P;
    signed char idxF[100], idxD[100];
};

struct Level {
    int64_t N = 0;     // nodes of this level (level 0: the owned nodes)
    int64_t Nvec = 0;  // nodes a vector of this level holds (level 0: including the halo, whose entries stay zero here)
    int nslices = 0;
    int64_t nnz_sell = 0;
    // matrix in DBSR / SELL-32 layout; level 0 aliases the handle's Jacobian
    const int32_t* sell_ptr = nullptr;
    const int32_t* colidx = nullptr;
    const double* offval = nullptr;
    const double* diagval = nullptr;
    const double* w = nullptr;  // symmetric edge weights in SELL order (strength graph)
    DevBuf<int32_t> sell_ptr_b, colidx_b;
    DevBuf<double> offval_b, diagval_b, w_b;
    DevBuf<double> binv;  // n*n planes x N: inverse diagonal blocks
    // transfer to the next coarser level
    int64_t Nc = 0;
    DevBuf<int32_t> agg;                 // N: aggregate of each node, -1 = not represented on the coarse level
    DevBuf<int32_t> agg_ptr, agg_nodes;  // Nc+1 / nodes sorted by aggregate
    int64_t nuniq = 0;
    DevBuf<int32_t> gal_ptr, gal_src, gal_dst;  // coarse entry u <- fine entries gal_src[gal_ptr[u] .. gal_ptr[u+1]); dst >= 0: SELL position, < 0: diagonal of node -dst-1
    // work vectors (n x Nvec); b2 / x2: second visit of a W-cycle
    DevBuf<double> x, b, t, b2, x2;
    // several ranks: halo of this level (level 0 uses the handle's lists)
    LevelHalo halo;
    std::vector<int32_t> send_idx_host;
    bool dist = false;  // rank-local rows of a distributed level: its SpMVs refresh the halo of their input (false: the level is complete on this rank)
};

// one captured V-/W-cycle per (input, output) vector pair: Krylov methods call the preconditioner with the same few pairs in every
// iteration, and the coarse levels are launch-latency bound (a W-cycle on cfg3 is 260 launches), so replaying a CUDA graph
// removes most of the launch overhead.  Single rank only: the peer exchanges carry a sequence number in their arguments.
struct CycleGraph {
    const double* in = nullptr;
    double* out = nullptr;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
};

struct Amg {
    std::vector<Level*> L;
    std::vector<CycleGraph> graphs;
    int applies = 0;  // applications since the last (re)build of the cycle
    // what a captured cycle bakes into its kernel arguments: a graph stays valid across numeric setups as long as none of it changes
    struct Signature {
        double omega = 0, alpha = 0;
        int sweeps = 0, coarse_sweeps = 0, wdepth = 0;
        const void *off0 = nullptr, *diag0 = nullptr, *off32 = nullptr;
        bool operator==(const Signature& o) const {
            return off32 == o.off32 && omega == o.omega && alpha == o.alpha && sweeps == o.sweeps && coarse_sweeps == o.coarse_sweeps && wdepth == o.wdepth && off0 == o.off0 && diag0 == o.diag0;
        }
    } captured;
    vfvm_handle* owner = nullptr;
    void drop_graphs() {
        for (CycleGraph& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
        applies = 0;
        if (owner) owner->graph_epoch++;  // a captured Krylov iteration contains this cycle
    }
    bool struct_valid = false, distributed = false;
    int64_t pattern_nnz = -1, pattern_N = -1;
    // Replicated coarse levels (several ranks).  A halo exchange costs ~15 us whatever its size, and a distributed cycle pays two of them
    // per level; below `repl_max_n` nodes (summed over the ranks) a level is pure latency, so from the first such level on every rank
    // holds the WHOLE level matrix and the whole sub-hierarchy under it: one all-gather of the restricted right-hand side per cycle
    // replaces all exchanges of those levels.  L[repl_level] is the replicated level (global numbering: rank r's nodes at r * repl_S,
    // its stored entries at r * repl_M); `local_part` is the same level as the distributed coarsening produced it (rank-local rows +
    // halo columns) and is only the target of the Galerkin product, whose values are all-gathered plane by plane in the numeric phase.
    int repl_level = 0;
    Level* local_part = nullptr;
    // The price is that every rank runs the replicated levels in full, so the rule is on WORK, not on nodes: a level is replicated when its
    // stored values (entries x planes, summed over the ranks) stay below repl_max_vals -- cfg3: from level 2 (63 k nodes, 1.6 M values), the
    // three-species system cfg4: from level 3 (7 k nodes, 1.6 M values; its level 2 holds 14 M values and costs more replicated on 8 GPUs
    // than its exchanges do: measured 3.5 against 2.6 ms per BiCGStab iteration, profiles/r2_amg_sweeps.txt).
    int64_t repl_S = 0, repl_M = 0, repl_max_n = 200000, repl_max_vals = 4000000;

    // fused coarse cycle (k_fused_cycle): levels >= fuse_level run in one persistent kernel
    bool fp32 = true;  // the SpMVs of the cycle on the finest level read an fp32 copy of the off-diagonal planes (VFVM_AMG_FP32=0: fp64)
    int fuse_level = 1;  // first level of the fused kernel (chosen by build_fused), 0 = off
    int64_t fuse_max_n = 131072;  // levels with more nodes than this keep their own (bandwidth-tuned) kernels
    bool fuse_valid = false;
    DevBuf<LevelDev> lev_dev;
    DevBuf<FuseOp> prog_dev[2];  // entry with the level's b / with its b2 (second visit of a W-cycle)
    int prog_len[2] = {0, 0};
    DevBuf<unsigned long long> bar;  // [0] barrier counter, [1] generation
    int fuse_grid = 0;
    double omega = 0.8, alpha = 1.75, theta = 0.08;  // alpha: measured on cfg3 (CG iterations 153 / 101 / 79 / 71 / 69 for alpha = 1 / 1.25 / 1.5 / 1.75 / 2)
    int coarse_sweeps = 4, max_levels = 20, sweeps = 1;
    int wdepth = 0;  // levels 1..wdepth are visited twice per visit of their parent (W-cycle on the top of the hierarchy), 0 = V-cycle  // sweeps: pre- and post-smoothing steps per level
    ~Amg() {
        drop_graphs();
        for (Level* l : L) delete l;
        delete local_part;
    }
};

// ------------------------------------------------------------------------------------------------ aggregation kernels
// All of them use the SpMV mapping: a warp owns a slice of 32 rows, one lane per row, entries e = base + 32 j + lane.
__device__ __forceinline__ bool strong(double w, double dK, double dL, double theta2) { return w > 0.0 && w * w >= theta2 * dK * dL; }
__device__ __forceinline__ unsigned long long prio_of(int K) {
    unsigned int x = (unsigned int)K * 2654435761u;
    x ^= x >> 15;
    x *= 2246822519u;
    x ^= x >> 13;
    return ((unsigned long long)x << 32) | (unsigned int)(K + 1);
}

#define ROW_LOOP_BEGIN(N_, nsl_)                                                                  \
    const int lane = threadIdx.x & 31;                                                            \
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);                            \
    if (g >= (nsl_)) return;                                                                      \
    const int K = g * 32 + lane;                                                                  \
    const bool valid = K < (N_);                                                                  \
    const int base = sell_ptr[g];                                                                 \
    const int wd = (sell_ptr[g + 1] - base) >> 5;

// node weight d_K = sum of the edge weights of the row (owned columns only); fixed nodes (whole diagonal block is a Dirichlet
// penalty) get d = inf, which makes every connection to them weak
template <int NS>
__global__ void k_node_weight(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                              const double* __restrict__ diagval, SpmvArgs a, int check_fixed, double* __restrict__ d) {
    ROW_LOOP_BEGIN(N, nsl)
    double s = 0.0;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (valid && L != K && L < N) s += w[e];
    }
    if (!valid) return;
    if (check_fixed) {
        bool all_fixed = true;
#pragma unroll
        for (int i = 0; i < NS; i++) all_fixed &= fabs(diagval[(int64_t)a.idxD[i * NS + i] * N + K]) >= 1.0e29;
        if (all_fixed) s = INFINITY;
    }
    d[K] = s;
}

// -1 undecided, -2 out (fixed or without strong neighbour)
__global__ void k_agg_init(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                           const double* __restrict__ d, double theta2, int32_t* __restrict__ agg) {
    ROW_LOOP_BEGIN(N, nsl)
    const double dK = valid ? d[K] : INFINITY;
    bool any = false;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (valid && L != K && L < N) any |= strong(w[e], dK, d[L], theta2);
    }
    if (valid) agg[K] = any ? -1 : -2;
}

// eligible = undecided and every strong neighbour undecided; ep = priority of eligible nodes, else 0
__global__ void k_agg_elig(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                           const double* __restrict__ d, double theta2, const int32_t* __restrict__ agg, unsigned long long* __restrict__ ep) {
    ROW_LOOP_BEGIN(N, nsl)
    const double dK = valid ? d[K] : INFINITY;
    bool ok = valid && agg[K] == -1;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (ok && L != K && L < N && strong(w[e], dK, d[L], theta2)) ok = agg[L] == -1;
    }
    if (valid) ep[K] = ok ? prio_of(K) : 0ull;
}

// out[K] = max(in[K], in[L] over strong neighbours L)
__global__ void k_agg_max(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                          const double* __restrict__ d, double theta2, const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out) {
    ROW_LOOP_BEGIN(N, nsl)
    const double dK = valid ? d[K] : INFINITY;
    unsigned long long m = valid ? in[K] : 0ull;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (valid && L != K && L < N && strong(w[e], dK, d[L], theta2)) m = max(m, in[L]);
    }
    if (valid) out[K] = m;
}

// an eligible node whose priority is the maximum within distance two becomes a root and takes its strong neighbours
__global__ void k_agg_root(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                           const double* __restrict__ d, double theta2, const unsigned long long* __restrict__ ep, const unsigned long long* __restrict__ m1,
                           int32_t* __restrict__ agg, int32_t* __restrict__ nroots) {
    ROW_LOOP_BEGIN(N, nsl)
    const double dK = valid ? d[K] : INFINITY;
    const unsigned long long mine = valid ? ep[K] : 0ull;
    unsigned long long m = valid ? m1[K] : 0ull;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (mine && L != K && L < N && strong(w[e], dK, d[L], theta2)) m = max(m, m1[L]);
    }
    const bool root = mine != 0ull && m == mine;
    if (root) {
        agg[K] = K;
        atomicAdd(nroots, 1);
    }
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (root && L != K && L < N && strong(w[e], dK, d[L], theta2)) agg[L] = K;  // no other root reaches L
    }
}

// undecided nodes join the aggregate of the neighbour they are most strongly tied to (reads `agg`, writes `out`)
__global__ void k_agg_join(int64_t N, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ w,
                           const double* __restrict__ d, double theta2, const int32_t* __restrict__ agg, int32_t* __restrict__ out) {
    ROW_LOOP_BEGIN(N, nsl)
    const double dK = valid ? d[K] : INFINITY;
    const int mine = valid ? agg[K] : -2;
    double best = 0.0;
    int target = -1;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        if (mine == -1 && L != K && L < N) {
            const double we = w[e];
            if (strong(we, dK, d[L], theta2) && we > best) {  // columns ascend: ties go to the lower node number
                const int aL = agg[L];
                if (aL >= 0) {
                    best = we;
                    target = aL;
                }
            }
        }
    }
    if (valid) out[K] = mine == -1 ? target : mine;  // target -1: still undecided
}

__global__ void k_agg_to_vec(int64_t N, int ns, const int32_t* __restrict__ agg, double* __restrict__ out) {  // aggregate id of each owned node as the first component of a level vector
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K < N) out[K * ns] = (double)agg[K];
}
__global__ void k_agg_rest(int64_t N, int32_t* __restrict__ agg, int32_t* __restrict__ isroot) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= N) return;
    if (agg[K] == -1) agg[K] = (int)K;  // nobody to join: an aggregate of its own
    isroot[K] = agg[K] == (int)K ? 1 : 0;
}
__global__ void k_agg_renumber(int64_t N, const int32_t* __restrict__ newid, int32_t* __restrict__ agg, int32_t* __restrict__ keys, int32_t* __restrict__ vals, int Nc) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= N) return;
    const int a = agg[K];
    const int id = a >= 0 ? newid[a] : -1;
    agg[K] = id;
    keys[K] = id >= 0 ? id : Nc;  // nodes outside the coarse problem sort to the end
    vals[K] = (int)K;
}
__global__ void k_lower_bounds32(int n, const int32_t* __restrict__ sorted, int64_t len, int32_t* __restrict__ out) {  // out[i] = first position with key >= i
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    out[i] = (int32_t)lo;
}

// ------------------------------------------------------------------------------------------------ coarse pattern kernels
#define KEY_SENTINEL 0xffffffffffffffffull
// Nc here = number of coarse COLUMNS (owned aggregates + coarse halo), the multiplier of the (row, column) key; columns up to
// Ncols_f (fine owned + halo nodes) take part, their aggregate ids come from `agg` (halo part filled by the owner's ids)
__global__ void k_pair_keys(int64_t N, int64_t Ncols_f, int nsl, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ agg, int64_t Nc,
                            unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
    ROW_LOOP_BEGIN(N, nsl)
    const int I = valid ? agg[K] : -1;
    for (int j = 0; j < wd; j++) {
        const int e = base + 32 * j + lane;
        const int L = colidx[e];
        unsigned long long key = KEY_SENTINEL;
        if (valid && I >= 0 && L != K && L < Ncols_f) {
            const int J = agg[L];
            if (J >= 0) key = (unsigned long long)I * (unsigned long long)Nc + (unsigned long long)J;
        }
        keys[e] = key;
        vals[e] = e;
    }
}
__global__ void k_flag_heads(int64_t n, const unsigned long long* __restrict__ keys, int32_t* __restrict__ flag) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flag[k] = (keys[k] != KEY_SENTINEL && (k == 0 || keys[k] != keys[k - 1])) ? 1 : 0;
}
__global__ void k_collect_uniques(int64_t n, const unsigned long long* __restrict__ keys, const int32_t* __restrict__ flag, const int32_t* __restrict__ uid,
                                  unsigned long long* __restrict__ ukey, int32_t* __restrict__ gal_ptr) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) {
        ukey[uid[k]] = keys[k];
        gal_ptr[uid[k]] = (int32_t)k;
    }
}
__device__ __forceinline__ int64_t lb64(const unsigned long long* __restrict__ a, int64_t n, unsigned long long v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__global__ void k_first_sentinel(const unsigned long long* __restrict__ a, int64_t n, int32_t* __restrict__ out) { *out = (int32_t)lb64(a, n, KEY_SENTINEL); }
// off-diagonal row lengths of the coarse matrix from the sorted unique (I,J) keys
__global__ void k_coarse_rowlen(int64_t Nrows, int64_t Nc, const unsigned long long* __restrict__ ukey, int64_t nuniq, int32_t* __restrict__ rowlen) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= Nrows) return;
    const int64_t f = lb64(ukey, nuniq, (unsigned long long)I * Nc), g = lb64(ukey, nuniq, (unsigned long long)(I + 1) * Nc);
    const unsigned long long dk = (unsigned long long)I * Nc + I;
    const int64_t dpos = lb64(ukey, nuniq, dk);
    const bool hasdiag = dpos < nuniq && ukey[dpos] == dk;
    rowlen[I] = (int32_t)(g - f - (hasdiag ? 1 : 0));
}
__global__ void k_slice_width(int nslc, int64_t Nc, const int32_t* __restrict__ rowlen, int32_t* __restrict__ width32) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= nslc) return;
    const int64_t I = (int64_t)g * 32 + lane;
    int m = I < Nc ? rowlen[I] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) width32[g] = 32 * m;
}
// padding entries point to their own row; the lanes of the last slice beyond the last row point to the last row (an index >= N
// would make the SpMV gather read past the end of the level vector)
__global__ void k_fill_rowid(int nslc, int64_t N, const int32_t* __restrict__ sell_ptr, int32_t* __restrict__ colidx) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= nslc) return;
    const int32_t self = (int32_t)min((int64_t)g * 32 + lane, N - 1);
    for (int e = sell_ptr[g] + lane; e < sell_ptr[g + 1]; e += 32) colidx[e] = self;
}
__global__ void k_place_uniques(int64_t nuniq, int64_t Nc, const unsigned long long* __restrict__ ukey, const int32_t* __restrict__ sell_ptr, int32_t* __restrict__ colidx,
                                int32_t* __restrict__ gal_dst) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nuniq) return;
    const unsigned long long key = ukey[u];
    const int64_t I = (int64_t)(key / (unsigned long long)Nc), J = (int64_t)(key - (unsigned long long)I * Nc);
    if (I == J) {
        gal_dst[u] = -(int32_t)I - 1;
        return;
    }
    const int64_t f = lb64(ukey, nuniq, (unsigned long long)I * Nc);
    int64_t j = u - f;
    if (J > I) {
        const unsigned long long dk = (unsigned long long)I * Nc + I;
        const int64_t dpos = lb64(ukey, nuniq, dk);
        if (dpos < nuniq && ukey[dpos] == dk) j--;
    }
    const int pos = sell_ptr[I >> 5] + 32 * (int)j + (int)(I & 31);
    colidx[pos] = (int32_t)J;
    gal_dst[u] = pos;
}

// ------------------------------------------------------------------------------------------------ numeric Galerkin product
// diagonal blocks: sum of the fine diagonal blocks of the aggregate (node list in ascending order)
__global__ void k_gal_diag(int64_t Nc, int cD, const int32_t* __restrict__ agg_ptr, const int32_t* __restrict__ agg_nodes, const double* __restrict__ fdiag, int64_t Nf,
                           double* __restrict__ cdiag) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= Nc) return;
    const int q0 = agg_ptr[I], q1 = agg_ptr[I + 1];
    for (int p = 0; p < cD; p++) {
        double s = 0.0;
        for (int q = q0; q < q1; q++) s += fdiag[(int64_t)p * Nf + agg_nodes[q]];
        cdiag[(int64_t)p * Nc + I] = s;
    }
}
// off-diagonal blocks (and the aggregate-internal fine blocks, which go to the coarse diagonal): gather in sorted order.
// planeD[p] = diagonal plane that holds the coupling of off-diagonal plane p.  One unique (I,I) per aggregate => single writer.
struct PlaneMap {
    int cF;
    int toD[100];
};
__global__ void k_gal_off(int64_t nuniq, PlaneMap pm, const int32_t* __restrict__ gal_ptr, const int32_t* __restrict__ gal_src, const int32_t* __restrict__ gal_dst,
                          const double* __restrict__ foff, int64_t fnnz, double* __restrict__ coff, int64_t cnnz, double* __restrict__ cdiag, int64_t Nc) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nuniq) return;
    const int k0 = gal_ptr[u], k1 = gal_ptr[u + 1], dst = gal_dst[u];
    for (int p = 0; p < pm.cF; p++) {
        double s = 0.0;
        for (int k = k0; k < k1; k++) s += foff[(int64_t)p * fnnz + gal_src[k]];
        if (dst >= 0) coff[(int64_t)p * cnnz + dst] = s;
        else cdiag[(int64_t)pm.toD[p] * Nc + (-dst - 1)] += s;
    }
}
__global__ void k_gal_weight(int64_t nuniq, const int32_t* __restrict__ gal_ptr, const int32_t* __restrict__ gal_src, const int32_t* __restrict__ gal_dst,
                             const double* __restrict__ fw, double* __restrict__ cw) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nuniq) return;
    const int dst = gal_dst[u];
    if (dst < 0) return;
    double s = 0.0;
    for (int k = gal_ptr[u]; k < gal_ptr[u + 1]; k++) s += fw[gal_src[k]];
    cw[dst] = s;
}

// ------------------------------------------------------------------------------------------------ cycle kernels
// x = omega B^-1 b  (first sweep from a zero guess)    or    x += omega B^-1 (b - t), t = A x; optionally mirrored to `out`
template <int NS>
__global__ void k_smooth(int64_t N, double omega, const double* __restrict__ binv, const double* __restrict__ b, const double* __restrict__ t, double* __restrict__ x,
                         double* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    double res[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) res[j] = t ? b[r * NS + j] - t[r * NS + j] : b[r * NS + j];
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++) s += binv[(int64_t)(i * NS + j) * N + r] * res[j];
        const double v = (t ? x[r * NS + i] : 0.0) + omega * s;
        x[r * NS + i] = v;
        if (out) out[r * NS + i] = v;
    }
}
// coarse right-hand side: b_c[I] = sum over the aggregate of (b - t), t = A x  (residual fused into the restriction)
template <int NS>
__global__ void k_restrict(int64_t Nc, const int32_t* __restrict__ agg_ptr, const int32_t* __restrict__ agg_nodes, const double* __restrict__ b, const double* __restrict__ t,
                           double* __restrict__ bc) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= Nc) return;
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) s[i] = 0.0;
    for (int q = agg_ptr[I]; q < agg_ptr[I + 1]; q++) {
        const int64_t K = agg_nodes[q];
#pragma unroll
        for (int i = 0; i < NS; i++) s[i] += b[K * NS + i] - t[K * NS + i];
    }
#pragma unroll
    for (int i = 0; i < NS; i++) bc[I * NS + i] = s[i];
}
// W-cycle helpers: r = b - t and x2 = x (before the second visit);  x += x2 (after it)
__global__ void k_w_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ t, const double* __restrict__ x, double* __restrict__ r, double* __restrict__ x2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = b[i] - t[i];
    x2[i] = x[i];
}
__global__ void k_w_add(int64_t n, const double* __restrict__ x2, double* __restrict__ x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += x2[i];
}
template <int NS>
__global__ void k_prolong(int64_t N, double alpha, const int32_t* __restrict__ agg, const double* __restrict__ xc, double* __restrict__ x) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= N) return;
    const int I = agg[K];
    if (I < 0) return;
#pragma unroll
    for (int i = 0; i < NS; i++) x[K * NS + i] += alpha * xc[(int64_t)I * NS + i];
}


// ------------------------------------------------------------------------------------------------ fused coarse cycle
// Everything below the finest level runs in ONE persistent kernel: the recursive cycle is unrolled by the host into a short program
// of level operations (smooth / SpMV / restrict / prolong / W-cycle bookkeeping), every block executes the program in lock step and
// a grid-wide barrier in global memory separates dependent operations (a sense-free counter barrier: ~1-2 us against ~2.5 us per
// kernel of a replayed graph and 4-5 us eager).  Levels that fit one block (N <= FUSE_SMALL_N) are processed by block 0 alone with
// __syncthreads between the operations, so the deepest levels of the hierarchy -- pure latency -- cost ~0.1 us per step.
// With several ranks the halo exchange of a level SpMV happens inside the kernel through the peer mailboxes (peer.cuh): push the
// boundary values, barrier, raise the flags, wait for the neighbours' flags, multiply.  The sequence counter is advanced in a
// register by every block alike and stored back once at the end of the kernel.
// A W-cycle on the coarse levels therefore costs no launches at all, which is what makes it affordable on every rank count.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long v;
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

template <int NS, bool DIAGMASK>
__device__ __forceinline__ void fuse_spmv(const FuseArgs& a, const LevelDev& l, const double* __restrict__ x, double* __restrict__ y, int warp, int nwarps, const double* hbox) {
    const int lane = threadIdx.x & 31;
    const int64_t nnz = l.nnz_sell;
    for (int g = warp; g < l.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < l.N;
        const int64_t r = valid ? rraw : l.N - 1;
        const int base = l.sell_ptr[g];
        const int w = (l.sell_ptr[g + 1] - base) >> 5;
        double acc[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) acc[i] = 0.0;
        constexpr int BATCH = NS == 1 ? 4 : 2;
        for (int j0 = 0; j0 < w; j0 += BATCH) {
            int Lc[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) Lc[b] = (j0 + b < w) ? l.colidx[(int64_t)base + (int64_t)(j0 + b) * 32 + lane] : (int)r;
            double xl[BATCH][NS];
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (hbox && Lc[b] >= l.N) {  // halo column: read from the mailbox (level lists -> mailbox position)
                    const int64_t c = Lc[b] - l.N;
                    int q = 0;
                    while (c >= l.recv_ptr[q + 1]) q++;
                    const int64_t pos = a.P.recv_ptr0[q] + (c - l.recv_ptr[q]);
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) xl[b][jj] = peer_ld_data(hbox + pos * NS + jj);
                } else {
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) xl[b][jj] = x[(int64_t)Lc[b] * NS + jj];
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (j0 + b >= w) break;
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
                if constexpr (DIAGMASK) {
#pragma unroll
                    for (int i = 0; i < NS; i++) acc[i] += l.offval[(int64_t)i * nnz + e] * xl[b][i];
                } else {
#pragma unroll
                    for (int i = 0; i < NS; i++)
#pragma unroll
                        for (int jj = 0; jj < NS; jj++) {
                            const int p = a.idxF[i * NS + jj];
                            if (p >= 0) acc[i] += l.offval[(int64_t)p * nnz + e] * xl[b][jj];
                        }
                }
            }
        }
        if (valid) {
            double xr[NS];
#pragma unroll
            for (int jj = 0; jj < NS; jj++) xr[jj] = x[r * NS + jj];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double sacc = acc[i];
                if constexpr (DIAGMASK) {
                    sacc += l.diagval[(int64_t)i * l.N + r] * xr[i];
                } else {
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) {
                        const int p = a.idxD[i * NS + jj];
                        if (p >= 0) sacc += l.diagval[(int64_t)p * l.N + r] * xr[jj];
                    }
                }
                y[r * NS + i] = sacc;
            }
        }
    }
}

template <int NS, bool DIAGMASK>
__global__ void __launch_bounds__(FUSE_THREADS) k_fused_cycle(const FuseArgs a) {
    const unsigned long long gen = *(volatile unsigned long long*)a.bar_gen;
    unsigned long long nbar = 0;
    unsigned long long seq = a.distributed ? *(volatile unsigned long long*)a.P.seq_ctr : 0ull;
    const int nblk = gridDim.x;
    for (int k = 0; k < a.nops; k++) {
        const FuseOp op = a.prog[k];
        const LevelDev& l = a.lev[op.lvl];
        const bool mine = !op.small || blockIdx.x == 0;
        // index space of this operation: the whole grid, or block 0 alone
        const int64_t tid = op.small ? threadIdx.x : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t nth = op.small ? blockDim.x : (int64_t)nblk * blockDim.x;
        const int warp = (int)(tid >> 5), nwarps = (int)(nth >> 5);
        const double* __restrict__ bvec = op.bsel ? l.b2 : l.b;
        const double* hbox = nullptr;
        if (op.code == FOP_SPMV && a.distributed) {
            // halo exchange of x_l through the mailboxes, inside the kernel
            seq++;
            const int par = (int)(seq & 1ull);
            if (mine) {
                const int64_t total = l.send_ptr[a.P.nn] * NS;
                for (int64_t i = tid; i < total; i += nth) {
                    const int64_t q = i / NS;
                    const int s = (int)(i - q * NS);
                    int r = 0;
                    while (q >= l.send_ptr[r + 1]) r++;
                    a.P.halo_dst[r][par * a.P.halo_dst_stride[r] + (q - l.send_ptr[r]) * NS + s] = l.x[(int64_t)l.send_idx[q] * NS + s];
                }
                __threadfence_system();
            }
            if (op.small) {
                if (mine) __syncthreads();
            } else {
                nbar++;
                grid_barrier(a.bar_ctr, gen + nbar * nblk);
            }
            if (blockIdx.x == 0 && threadIdx.x < a.P.nn) peer_st_flag(a.P.hflag_dst[threadIdx.x] + par * a.P.nranks, seq);
            if (mine) {
                if (threadIdx.x < a.P.nn) peer_wait(a.P.hflag_local + par * a.P.nranks + threadIdx.x, seq, a.P.err, a.P.timeout_ns);
                __syncthreads();
            }
            hbox = a.P.halo_local + (int64_t)par * a.P.halo_local_stride;
        }
        if (mine) {
            switch (op.code) {
                case FOP_SMOOTH_FIRST:
                case FOP_SMOOTH: {
                    const bool first = op.code == FOP_SMOOTH_FIRST;
                    for (int64_t r = tid; r < l.N; r += nth) {
                        double res[NS];
#pragma unroll
                        for (int j = 0; j < NS; j++) res[j] = first ? bvec[r * NS + j] : bvec[r * NS + j] - l.t[r * NS + j];
#pragma unroll
                        for (int i = 0; i < NS; i++) {
                            double sacc = 0.0;
#pragma unroll
                            for (int j = 0; j < NS; j++) sacc += l.binv[(int64_t)(i * NS + j) * l.N + r] * res[j];
                            l.x[r * NS + i] = (first ? 0.0 : l.x[r * NS + i]) + a.omega * sacc;
                        }
                    }
                    break;
                }
                case FOP_SPMV: fuse_spmv<NS, DIAGMASK>(a, l, l.x, l.t, warp, nwarps, hbox); break;
                case FOP_RESTRICT: {
                    const LevelDev& c = a.lev[op.lvl + 1];
                    for (int64_t I = tid; I < c.N; I += nth) {
                        double sacc[NS];
#pragma unroll
                        for (int i = 0; i < NS; i++) sacc[i] = 0.0;
                        for (int q = l.agg_ptr[I]; q < l.agg_ptr[I + 1]; q++) {
                            const int64_t K = l.agg_nodes[q];
#pragma unroll
                            for (int i = 0; i < NS; i++) sacc[i] += bvec[K * NS + i] - l.t[K * NS + i];
                        }
#pragma unroll
                        for (int i = 0; i < NS; i++) c.b[I * NS + i] = sacc[i];
                    }
                    break;
                }
                case FOP_WRES:
                    for (int64_t i = tid; i < l.N * NS; i += nth) {
                        l.b2[i] = l.b[i] - l.t[i];
                        l.x2[i] = l.x[i];
                    }
                    break;
                case FOP_WADD:
                    for (int64_t i = tid; i < l.N * NS; i += nth) l.x[i] += l.x2[i];
                    break;
                case FOP_PROLONG: {
                    const LevelDev& c = a.lev[op.lvl + 1];
                    for (int64_t K = tid; K < l.N; K += nth) {
                        const int I = l.agg[K];
                        if (I < 0) continue;
#pragma unroll
                        for (int i = 0; i < NS; i++) l.x[K * NS + i] += a.alpha * c.x[(int64_t)I * NS + i];
                    }
                    break;
                }
                default: break;
            }
        }
        // separation from the next operation: block-local if both run on block 0 alone, grid-wide otherwise
        if (k + 1 < a.nops) {
            if (op.small && a.prog[k + 1].small) {
                if (mine) __syncthreads();
            } else {
                nbar++;
                grid_barrier(a.bar_ctr, gen + nbar * nblk);
            }
        }
    }
    nbar++;
    grid_barrier(a.bar_ctr, gen + nbar * nblk);  // every block has read `gen` (and the sequence counter) by now
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.bar_gen = gen + nbar * nblk;
        if (a.distributed) *a.P.seq_ctr = seq;
    }
}

#define NS_SWITCH(n, ...)                                                                  \
    switch (n) {                                                                           \
        case 1: { constexpr int NS = 1; __VA_ARGS__; } break;                              \
        case 2: { constexpr int NS = 2; __VA_ARGS__; } break;                              \
        case 3: { constexpr int NS = 3; __VA_ARGS__; } break;                              \
        case 4: { constexpr int NS = 4; __VA_ARGS__; } break;                              \
        case 5: { constexpr int NS = 5; __VA_ARGS__; } break;                              \
        case 10: { constexpr int NS = 10; __VA_ARGS__; } break;                            \
        default: throw std::string("number of species without device instantiation");      \
    }

SpmvArgs level_args(vfvm_handle* h, const Level& l) {
    SpmvArgs a = vfvm_spmv_args(h);  // plane tables
    a.sell_ptr = l.sell_ptr;
    a.colidx = l.colidx;
    a.offval = l.offval;
    a.diagval = l.diagval;
    a.nnz_sell = l.nnz_sell;
    a.Nown = l.N;
    a.nslices = l.nslices;
    return a;
}

template <class T>
T fetch(const T* dev, cudaStream_t s) {
    T v;
    CK(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return v;
}

// aggregates of level l; returns the number of aggregates
int64_t aggregate(vfvm_handle* h, Amg& A, Level& l, bool level0) {
    cudaStream_t s = h->stream;
    const int64_t N = l.N;
    const int nsl = l.nslices, gridw = cdiv(nsl, 8);
    const double th2 = A.theta * A.theta;
    DevBuf<double> d;
    d.alloc(N);
    SpmvArgs a = level_args(h, l);
    NS_SWITCH(h->n, (k_node_weight<NS><<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, l.diagval, a, level0 ? 1 : 0, d.p)));
    l.agg.alloc(l.Nvec);
    CK(cudaMemsetAsync(l.agg.p, 0xff, l.Nvec * sizeof(int32_t), s));  // halo part: -1 until the owners' ids arrive
    k_agg_init<<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, d.p, th2, l.agg.p);
    DevBuf<unsigned long long> ep, m1;
    ep.alloc(N);
    m1.alloc(N);
    DevBuf<int32_t> cnt, agg2, isroot, newid;
    cnt.alloc(1);
    h->launches += 2;
    for (int round = 0; round < 64; round++) {
        CK(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t), s));
        k_agg_elig<<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, d.p, th2, l.agg.p, ep.p);
        k_agg_max<<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, d.p, th2, ep.p, m1.p);
        k_agg_root<<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, d.p, th2, ep.p, m1.p, l.agg.p, cnt.p);
        h->launches += 3;
        if (fetch(cnt.p, s) == 0) break;
    }
    agg2.alloc(N);
    for (int pass = 0; pass < 2; pass++) {  // left-overs join a neighbouring aggregate; the second pass serves nodes whose neighbours joined in the first
        k_agg_join<<<gridw, 256, 0, s>>>(N, nsl, l.sell_ptr, l.colidx, l.w, d.p, th2, l.agg.p, agg2.p);
        CK(cudaMemcpyAsync(l.agg.p, agg2.p, N * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        h->launches++;
    }
    isroot.alloc(N);
    newid.alloc(N);
    k_agg_rest<<<cdiv(N, 256), 256, 0, s>>>(N, l.agg.p, isroot.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, isroot.p, newid.p, (int)N, s);
    DevBuf<char> tmp;
    tmp.alloc(tb + 16);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, isroot.p, newid.p, (int)N, s);
    const int64_t Nc = (int64_t)fetch(newid.p + (N - 1), s) + fetch(isroot.p + (N - 1), s);
    // node lists of the aggregates
    DevBuf<int32_t> keys, vals, keys_s;
    keys.alloc(N);
    vals.alloc(N);
    keys_s.alloc(N);
    l.agg_nodes.alloc(N);
    k_agg_renumber<<<cdiv(N, 256), 256, 0, s>>>(N, newid.p, l.agg.p, keys.p, vals.p, (int)Nc);
    int bits = 1;
    while ((1ll << bits) <= Nc) bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys_s.p, vals.p, l.agg_nodes.p, (int)N, 0, bits, s);
    tmp.alloc(tb + 16);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_s.p, vals.p, l.agg_nodes.p, (int)N, 0, bits, s);
    l.agg_ptr.alloc(Nc + 1);
    k_lower_bounds32<<<cdiv(Nc + 1, 256), 256, 0, s>>>((int)Nc + 1, keys_s.p, N, l.agg_ptr.p);
    h->launches += 6;
    CK(cudaStreamSynchronize(s));
    l.Nc = Nc;
    return Nc;
}

// pattern of the next level + gather maps of the Galerkin product
void coarsen_pattern(vfvm_handle* h, Level& f, Level& c) {
    cudaStream_t s = h->stream;
    const int64_t Nrows = f.Nc, nnz = f.nnz_sell;
    c.N = Nrows;
    c.Nvec = Nrows + c.halo.nhalo;   // owned aggregates + coarse halo
    const int64_t Nc = c.Nvec;        // number of coarse columns = multiplier of the (row, column) keys
    c.nslices = cdiv(Nrows, 32);
    DevBuf<unsigned long long> keys, keys_s, ukey;
    DevBuf<int32_t> vals, flag, uid;
    keys.alloc(nnz);
    keys_s.alloc(nnz);
    vals.alloc(nnz);
    f.gal_src.alloc(nnz);
    k_pair_keys<<<cdiv(f.nslices, 8), 256, 0, s>>>(f.N, f.Nvec, f.nslices, f.sell_ptr, f.colidx, f.agg.p, Nc, keys.p, vals.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys_s.p, vals.p, f.gal_src.p, (int)nnz, 0, 64, s);
    DevBuf<char> tmp;
    tmp.alloc(tb + 16);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_s.p, vals.p, f.gal_src.p, (int)nnz, 0, 64, s);
    keys.release();
    vals.release();
    flag.alloc(nnz);
    uid.alloc(nnz);
    k_flag_heads<<<cdiv(nnz, 256), 256, 0, s>>>(nnz, keys_s.p, flag.p);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, uid.p, (int)nnz, s);
    tmp.alloc(tb + 16);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, uid.p, (int)nnz, s);
    const int64_t nuniq = (int64_t)fetch(uid.p + (nnz - 1), s) + fetch(flag.p + (nnz - 1), s);
    f.nuniq = nuniq;
    ukey.alloc(std::max<int64_t>(1, nuniq));
    f.gal_ptr.alloc(nuniq + 1);
    f.gal_dst.alloc(std::max<int64_t>(1, nuniq));
    k_collect_uniques<<<cdiv(nnz, 256), 256, 0, s>>>(nnz, keys_s.p, flag.p, uid.p, ukey.p, f.gal_ptr.p);
    // number of valid (non-sentinel) sorted entries = end of the last segment
    {
        DevBuf<int32_t> nv;
        nv.alloc(1);
        k_first_sentinel<<<1, 1, 0, s>>>(keys_s.p, nnz, nv.p);
        CK(cudaMemcpyAsync(f.gal_ptr.p + nuniq, nv.p, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    DevBuf<int32_t> rowlen, width32;
    rowlen.alloc(std::max<int64_t>(1, Nrows));
    width32.alloc(c.nslices + 1);
    CK(cudaMemsetAsync(width32.p, 0, (c.nslices + 1) * sizeof(int32_t), s));
    if (Nrows) {
        k_coarse_rowlen<<<cdiv(Nrows, 256), 256, 0, s>>>(Nrows, Nc, ukey.p, nuniq, rowlen.p);
        k_slice_width<<<cdiv(c.nslices, 8), 256, 0, s>>>(c.nslices, Nrows, rowlen.p, width32.p);
    }
    c.sell_ptr_b.alloc(c.nslices + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, width32.p, c.sell_ptr_b.p, c.nslices + 1, s);
    tmp.alloc(tb + 16);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, width32.p, c.sell_ptr_b.p, c.nslices + 1, s);
    c.nnz_sell = fetch(c.sell_ptr_b.p + c.nslices, s);
    c.colidx_b.alloc(std::max<int64_t>(1, c.nnz_sell));
    c.w_b.alloc(std::max<int64_t>(1, c.nnz_sell));
    CK(cudaMemsetAsync(c.w_b.p, 0, std::max<int64_t>(1, c.nnz_sell) * sizeof(double), s));
    if (c.nslices) k_fill_rowid<<<cdiv(c.nslices, 8), 256, 0, s>>>(c.nslices, Nrows, c.sell_ptr_b.p, c.colidx_b.p);
    if (nuniq) {
        k_place_uniques<<<cdiv(nuniq, 256), 256, 0, s>>>(nuniq, Nc, ukey.p, c.sell_ptr_b.p, c.colidx_b.p, f.gal_dst.p);
        k_gal_weight<<<cdiv(nuniq, 256), 256, 0, s>>>(nuniq, f.gal_ptr.p, f.gal_src.p, f.gal_dst.p, f.w, c.w_b.p);
    }
    h->launches += 10;
    c.offval_b.alloc((size_t)std::max(1, h->cF) * std::max<int64_t>(1, c.nnz_sell));
    CK(cudaMemsetAsync(c.offval_b.p, 0, c.offval_b.n * sizeof(double), s));  // padding entries stay exact zeros
    c.diagval_b.alloc((size_t)std::max(1, h->cD) * std::max<int64_t>(1, Nrows));
    c.sell_ptr = c.sell_ptr_b.p;
    c.colidx = c.colidx_b.p;
    c.offval = c.offval_b.p;
    c.diagval = c.diagval_b.p;
    c.w = c.w_b.p;
    CK(cudaStreamSynchronize(s));
}

// sum over the ranks of a host value (collective; identity on one rank)
double global_sum(vfvm_handle* h, double v) {
    if (h->nranks <= 1) return v;
    DevBuf<double> d;
    d.alloc(1);
    CK(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    vfvm_comm_allreduce_sum(h, d.p, 1);
    return fetch(d.p, h->stream);
}

void exchange(vfvm_handle* h, Amg& A, size_t i, double* x) {  // halo refresh of a level vector
    if (!A.distributed) return;
    if (i == 0) vfvm_halo_exchange_ptr(h, x);
    else vfvm_halo_exchange_level(h, A.L[i]->halo, x);
}

double global_max(vfvm_handle* h, double v) {
    if (h->nranks <= 1) return v;
    DevBuf<double> d;
    d.alloc(1);
    CK(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    vfvm_comm_allreduce_max(h, d.p, 1);
    return fetch(d.p, h->stream);
}

// ---- replicated coarse levels (see Amg::repl_level) -------------------------------------------------------------------------------
__global__ void k_iota_vec(int64_t N, int ns, double* __restrict__ out) {  // local node id as the first component of a level vector
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K < N) out[K * ns] = (double)K;
}
// my segment of the replicated pattern: local entries with their columns in global numbering; the rest of the segment belongs to the
// rank's last (padding) slice, whose rows point to themselves with weight zero
__global__ void k_repl_fill(int64_t nnz_loc, int64_t M, int64_t seg0, int64_t padrow0, const int32_t* __restrict__ colidx, const int32_t* __restrict__ colmap,
                            const double* __restrict__ w, int32_t* __restrict__ colidx_g, double* __restrict__ w_g) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    if (e < nnz_loc) {
        colidx_g[seg0 + e] = colmap[colidx[e]];
        w_g[seg0 + e] = w[e];
    } else {
        colidx_g[seg0 + e] = (int32_t)(padrow0 + (e & 31));
        w_g[seg0 + e] = 0.0;
    }
}

// Builds the replicated level G from the distributed level c (rank-local rows, halo columns).  Collective.
// Layout: every rank gets S rows (a multiple of 32 with at least one whole padding slice) and M stored entries; rank r's slices keep
// their local SELL order at entry offset r * M, and its last padding slice absorbs the unused part of the segment, so that slice g of the
// global matrix still ends where slice g + 1 begins.  Padding rows are identity rows (diagonal 1, right-hand side 0).  With equal
// segment sizes every plane is all-gathered in place.
void replicate_level(vfvm_handle* h, Amg& A, Level& c, Level& G) {
    cudaStream_t s = h->stream;
    const int n = h->n, R = h->nranks, rank = h->rank, nn = (int)h->nb_ranks.size();
    const int64_t Nloc = c.N, nnz_loc = c.nnz_sell;
    const int64_t S = ((int64_t)global_max(h, (double)Nloc) + 31) / 32 * 32 + 32;
    const int64_t M = std::max<int64_t>(32, ((int64_t)global_max(h, (double)nnz_loc) + 31) / 32 * 32);
    A.repl_S = S;
    A.repl_M = M;
    // global ids of my columns: owned = rank * S + id; halo = owner's rank * S + the owner's local id (one halo exchange of the ids)
    std::vector<int32_t> colmap((size_t)std::max<int64_t>(1, c.Nvec), 0);
    for (int64_t K = 0; K < Nloc; K++) colmap[K] = (int32_t)(rank * S + K);
    {
        DevBuf<double> T;
        T.alloc((size_t)n * std::max<int64_t>(1, c.Nvec));
        CK(cudaMemsetAsync(T.p, 0, T.n * sizeof(double), s));
        if (Nloc) k_iota_vec<<<cdiv(Nloc, 256), 256, 0, s>>>(Nloc, n, T.p);
        vfvm_halo_exchange_level(h, c.halo, T.p);
        std::vector<double> Th = T.to_host(s);
        for (int r = 0; r < nn; r++)
            for (int64_t q = c.halo.recv_ptr[r]; q < c.halo.recv_ptr[r + 1]; q++)
                colmap[Nloc + q] = (int32_t)((int64_t)h->nb_ranks[r] * S + (int64_t)Th[(size_t)(Nloc + q) * n]);
    }
    G.N = G.Nvec = (int64_t)R * S;
    G.nslices = (int)(G.N / 32);
    G.nnz_sell = (int64_t)R * M;
    G.dist = false;
    G.sell_ptr_b.alloc(G.nslices + 1);
    G.colidx_b.alloc(G.nnz_sell);
    G.w_b.alloc(G.nnz_sell);
    G.offval_b.alloc((size_t)std::max(1, h->cF) * G.nnz_sell);
    G.diagval_b.alloc((size_t)std::max(1, h->cD) * G.N);
    CK(cudaMemsetAsync(G.offval_b.p, 0, G.offval_b.n * sizeof(double), s));  // padding entries and segment tails stay exact zeros
    CK(cudaMemsetAsync(G.diagval_b.p, 0, G.diagval_b.n * sizeof(double), s));
    DevBuf<int32_t> cm;
    cm.upload(colmap.data(), colmap.size(), s);
    k_repl_fill<<<cdiv(M, 256), 256, 0, s>>>(nnz_loc, M, (int64_t)rank * M, (int64_t)rank * S + S - 32, c.colidx, cm.p, c.w, G.colidx_b.p, G.w_b.p);
    h->launches += 2;
    vfvm_comm_allgather_bytes(h, G.colidx_b.p, (size_t)M * sizeof(int32_t));
    vfvm_comm_allgather_bytes(h, G.w_b.p, (size_t)M * sizeof(double));
    // slice pointers: local ones shifted by the segment offset; slices past the local ones are empty except the last, which ends at the
    // next rank's segment
    const int spr = (int)(S / 32);
    std::vector<int32_t> sp_loc = c.sell_ptr_b.to_host(s);
    std::vector<int32_t> seg((size_t)spr);
    for (int g = 0; g < spr; g++) seg[g] = g < c.nslices ? sp_loc[g] : (int32_t)nnz_loc;
    DevBuf<int32_t> segs;
    segs.alloc((size_t)R * spr);
    CK(cudaMemcpyAsync(segs.p + (size_t)rank * spr, seg.data(), (size_t)spr * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    vfvm_comm_allgather_bytes(h, segs.p, (size_t)spr * sizeof(int32_t));
    std::vector<int32_t> all = segs.to_host(s);
    std::vector<int32_t> spg((size_t)G.nslices + 1);
    for (int r = 0; r < R; r++)
        for (int g = 0; g < spr; g++) spg[(size_t)r * spr + g] = (int32_t)((int64_t)r * M + all[(size_t)r * spr + g]);
    spg[(size_t)G.nslices] = (int32_t)((int64_t)R * M);
    G.sell_ptr_b.upload(spg.data(), spg.size(), s);
    // identity diagonal of my padding rows (the other ranks' arrive with the all-gather of the numeric phase)
    std::vector<double> ones((size_t)(S - Nloc), 1.0);
    for (int i = 0; i < n; i++) {
        const int p = h->idxD[i * n + i];
        if (p >= 0) CK(cudaMemcpyAsync(G.diagval_b.p + (size_t)p * G.N + (size_t)rank * S + Nloc, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    CK(cudaStreamSynchronize(s));
    G.sell_ptr = G.sell_ptr_b.p;
    G.colidx = G.colidx_b.p;
    G.offval = G.offval_b.p;
    G.diagval = G.diagval_b.p;
    G.w = G.w_b.p;
    vfvm_gather_box_create(h, S * n);
    h->amg_nccl_in_cycle = !h->gbox_ok;
}

// numeric phase: my rows of the replicated level come from the Galerkin product into the local part; every plane is then all-gathered in
// place (equal segment sizes).  NCCL: the volume is the whole level matrix (cfg4 at 193^3: 9 planes x 1.7 M entries), bandwidth bound.
void replicate_values(vfvm_handle* h, Amg& A) {
    cudaStream_t s = h->stream;
    Level &c = *A.local_part, &G = *A.L[A.repl_level];
    const int64_t S = A.repl_S, M = A.repl_M, rank = h->rank;
    for (int p = 0; p < h->cF && c.nnz_sell; p++)
        CK(cudaMemcpyAsync(G.offval_b.p + (size_t)p * G.nnz_sell + (size_t)rank * M, c.offval_b.p + (size_t)p * c.nnz_sell, (size_t)c.nnz_sell * sizeof(double), cudaMemcpyDeviceToDevice, s));
    for (int p = 0; p < h->cD && c.N; p++)
        CK(cudaMemcpyAsync(G.diagval_b.p + (size_t)p * G.N + (size_t)rank * S, c.diagval_b.p + (size_t)p * c.N, (size_t)c.N * sizeof(double), cudaMemcpyDeviceToDevice, s));
    vfvm_comm_group_start();
    for (int p = 0; p < h->cF; p++) vfvm_comm_allgather_bytes(h, G.offval_b.p + (size_t)p * G.nnz_sell, (size_t)M * sizeof(double));
    for (int p = 0; p < h->cD; p++) vfvm_comm_allgather_bytes(h, G.diagval_b.p + (size_t)p * G.N, (size_t)S * sizeof(double));
    vfvm_comm_group_end();
}

// Halo of the next level (several ranks).  Every rank sends the aggregate ids of its boundary nodes to the neighbours (one halo
// exchange on level f); the distinct ids per neighbour, ascending, are the coarse halo nodes, and the owner builds the matching
// send list from the same ids -- both sides see the same multiset in the same order, so no second exchange is needed.
void build_coarse_halo(vfvm_handle* h, Amg& A, size_t fi, Level& f, Level& c) {
    cudaStream_t s = h->stream;
    const int n = h->n, nn = (int)h->nb_ranks.size();
    const std::vector<int64_t>& fsp = fi == 0 ? h->send_ptr : f.halo.send_ptr;
    const std::vector<int64_t>& frp = fi == 0 ? h->recv_ptr : f.halo.recv_ptr;
    if (fi == 0) f.send_idx_host = h->send_idx.to_host(s);
    const int64_t nhalo_f = f.Nvec - f.N;
    DevBuf<double> T;
    T.alloc((size_t)n * f.Nvec);
    CK(cudaMemsetAsync(T.p, 0, T.n * sizeof(double), s));
    if (f.N) k_agg_to_vec<<<cdiv(f.N, 256), 256, 0, s>>>(f.N, n, f.agg.p, T.p);
    exchange(h, A, fi, T.p);
    std::vector<double> Th = T.to_host(s);
    std::vector<int32_t> aggh = f.agg.to_host(s);
    // receiver side
    std::vector<int32_t> halo_agg((size_t)std::max<int64_t>(1, nhalo_f), -1);
    c.halo.recv_ptr.assign(nn + 1, 0);
    for (int r = 0; r < nn; r++) {
        std::vector<int32_t> ids;
        for (int64_t q = frp[r]; q < frp[r + 1]; q++) {
            const int32_t a = (int32_t)Th[(size_t)(f.N + q) * n];
            if (a >= 0) ids.push_back(a);
        }
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        for (int64_t q = frp[r]; q < frp[r + 1]; q++) {
            const int32_t a = (int32_t)Th[(size_t)(f.N + q) * n];
            if (a >= 0) halo_agg[q] = (int32_t)(f.Nc + c.halo.recv_ptr[r] + (std::lower_bound(ids.begin(), ids.end(), a) - ids.begin()));
        }
        c.halo.recv_ptr[r + 1] = c.halo.recv_ptr[r] + (int64_t)ids.size();
    }
    if (nhalo_f) CK(cudaMemcpyAsync(f.agg.p + f.N, halo_agg.data(), nhalo_f * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // owner side
    c.halo.send_ptr.assign(nn + 1, 0);
    c.send_idx_host.clear();
    for (int r = 0; r < nn; r++) {
        std::vector<int32_t> ids;
        for (int64_t q = fsp[r]; q < fsp[r + 1]; q++) {
            const int32_t a = aggh[f.send_idx_host[q]];
            if (a >= 0) ids.push_back(a);
        }
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        c.send_idx_host.insert(c.send_idx_host.end(), ids.begin(), ids.end());
        c.halo.send_ptr[r + 1] = (int64_t)c.send_idx_host.size();
    }
    c.halo.Nown = f.Nc;
    c.halo.nhalo = c.halo.recv_ptr[nn];
    c.halo.send_idx.upload(c.send_idx_host.data(), c.send_idx_host.size(), s);
    CK(cudaStreamSynchronize(s));
}

void build_hierarchy(vfvm_handle* h, Amg& A) {
    A.drop_graphs();
    A.fuse_valid = false;
    for (Level* l : A.L) delete l;
    A.L.clear();
    delete A.local_part;
    A.local_part = nullptr;
    A.repl_level = 0;
    h->amg_nccl_in_cycle = false;
    A.distributed = h->nranks > 1 && !h->nb_ranks.empty() && !getenv("VFVM_AMG_LOCAL");
    if (const char* e = getenv("VFVM_AMG_REPL_MAX_N")) A.repl_max_n = atoll(e);
    if (const char* e = getenv("VFVM_AMG_REPL_MAX_VALS")) A.repl_max_vals = atoll(e);
    Level* l0 = new Level();
    l0->N = h->Nown;
    l0->Nvec = h->N;
    l0->nslices = h->ngroups;
    l0->nnz_sell = h->nnz_sell;
    l0->sell_ptr = h->sell_ptr.p;
    l0->colidx = h->colidx.p;
    l0->offval = h->offval.p;
    l0->diagval = h->diagval.p;
    l0->w = h->nzfac.p;
    l0->dist = A.distributed;
    A.L.push_back(l0);
    // below the replicated level every rank holds the same complete levels and takes the same decisions without communication
    auto gsum = [&](double v) { return A.repl_level ? v : global_sum(h, v); };
    while ((int)A.L.size() < A.max_levels) {
        Level& f = *A.L.back();
        // every decision about the depth is taken on rank-summed numbers: all ranks must build the same number of levels
        const double Nglob = gsum((double)f.N);
        if (Nglob <= 64.0 * (A.repl_level ? 1 : h->nranks)) break;
        if (gsum(f.N < 16 ? 1.0 : 0.0) > 0.0) break;  // some rank has (almost) run out of nodes: this level is the coarsest everywhere
        const int64_t Nc = f.N > 0 ? aggregate(h, A, f, A.L.size() == 1) : 0;
        if (f.N == 0) {
            f.agg.alloc(std::max<int64_t>(1, f.Nvec));
            f.agg_ptr.alloc(1);
            CK(cudaMemsetAsync(f.agg_ptr.p, 0, sizeof(int32_t), h->stream));
            f.Nc = 0;
        }
        const double Ncglob = gsum((double)Nc);
        if (Ncglob < 1.0 || Ncglob > 0.8 * Nglob) {  // no real coarsening any more: this level is the coarsest
            f.Nc = 0;
            break;
        }
        Level* c = new Level();
        const bool dist_level = A.distributed && !A.repl_level;
        if (dist_level) build_coarse_halo(h, A, A.L.size() - 1, f, *c);
        coarsen_pattern(h, f, *c);
        c->dist = dist_level;
        const double vals_glob = dist_level ? gsum((double)c->nnz_sell * std::max(1, h->cF)) : 0.0;
        if (dist_level && A.repl_max_n > 0 && Ncglob <= (double)A.repl_max_n && vals_glob <= (double)A.repl_max_vals) {  // small enough: replicate this level and everything below
            Level* G = new Level();
            replicate_level(h, A, *c, *G);
            A.local_part = c;
            A.repl_level = (int)A.L.size();
            A.L.push_back(G);
        } else {
            A.L.push_back(c);
        }
    }
    A.L.back()->Nc = 0;
    const int n = h->n;
    for (size_t i = 0; i < A.L.size(); i++) {
        Level& l = *A.L[i];
        l.binv.alloc((size_t)n * n * std::max<int64_t>(1, l.N));
        l.x.alloc((size_t)n * std::max<int64_t>(1, l.Nvec));
        l.t.alloc((size_t)n * std::max<int64_t>(1, l.Nvec));
        CK(cudaMemsetAsync(l.x.p, 0, l.x.n * sizeof(double), h->stream));
        CK(cudaMemsetAsync(l.t.p, 0, l.t.n * sizeof(double), h->stream));
        if (i > 0) {
            l.b.alloc((size_t)n * std::max<int64_t>(1, l.Nvec));
            l.b2.alloc((size_t)n * std::max<int64_t>(1, l.Nvec));
            l.x2.alloc((size_t)n * std::max<int64_t>(1, l.Nvec));
            CK(cudaMemsetAsync(l.b.p, 0, l.b.n * sizeof(double), h->stream));  // (padding rows of a replicated level keep a zero right-hand side)
            CK(cudaMemsetAsync(l.b2.p, 0, l.b2.n * sizeof(double), h->stream));
            CK(cudaMemsetAsync(l.x2.p, 0, l.x2.n * sizeof(double), h->stream));
        }
    }
    A.struct_valid = true;
    A.pattern_nnz = h->nnz_sell;
    A.pattern_N = h->Nown;
    if (getenv("VFVM_AMG_VERBOSE")) {
        fprintf(stderr, "[vfvm amg] rank %d %s levels (nodes+halo(stored blocks)):", h->rank, A.distributed ? "distributed" : "local");
        for (size_t i = 0; i < A.L.size(); i++) {
            Level* l = A.L[i];
            fprintf(stderr, " %s%lld+%lld(%lld)", A.repl_level && (int)i == A.repl_level ? "| replicated: " : "", (long long)l->N, (long long)(l->Nvec - l->N), (long long)l->nnz_sell);
        }
        fprintf(stderr, "\n");
    }
}

__global__ void k_to_float(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

void numeric_setup(vfvm_handle* h, Amg& A) {
    cudaStream_t s = h->stream;
    // Preconditioner-internal precision: the two finest-level SpMVs of a cycle are its bandwidth; their off-diagonal planes are read from an
    // fp32 copy (the diagonal blocks, the vectors and the accumulation stay fp64, and so does the outer Krylov method with the fp64 matrix --
    // the solution converges to the same tolerance, only the preconditioner is perturbed at the 1e-7 level)
    if (const char* e = getenv("VFVM_AMG_FP32")) A.fp32 = atoi(e) != 0;
    if (A.fp32 && h->cF > 0) {
        const int64_t nv = (int64_t)h->cF * h->nnz_sell;
        h->offval32.alloc((size_t)nv);
        k_to_float<<<148 * 8, 256, 0, s>>>(nv, h->offval.p, h->offval32.p);
        h->launches++;
    } else {
        h->offval32.release();
    }
    PlaneMap pm;
    pm.cF = h->cF;
    for (int p = 0; p < 100; p++) pm.toD[p] = p < h->cF ? h->idxD[h->planeF[p]] : 0;
    A.L[0]->offval = h->offval.p;  // (the handle may have reallocated its planes)
    A.L[0]->diagval = h->diagval.p;
    for (size_t i = 0; i + 1 < A.L.size(); i++) {
        const bool to_repl = A.repl_level && (int)(i + 1) == A.repl_level;  // the product lands in the rank-local part, then travels
        Level &f = *A.L[i], &c = to_repl ? *A.local_part : *A.L[i + 1];
        if (c.N) k_gal_diag<<<cdiv(c.N, 128), 128, 0, s>>>(c.N, h->cD, f.agg_ptr.p, f.agg_nodes.p, f.diagval, f.N, c.diagval_b.p);
        if (f.nuniq)
            k_gal_off<<<cdiv(f.nuniq, 128), 128, 0, s>>>(f.nuniq, pm, f.gal_ptr.p, f.gal_src.p, f.gal_dst.p, f.offval, f.nnz_sell, c.offval_b.p, c.nnz_sell, c.diagval_b.p, c.N);
        h->launches += 2;
        if (to_repl) replicate_values(h, A);
    }
    for (Level* l : A.L)
        if (l->N) vfvm_blockinv_level(h, level_args(h, *l), l->N, l->diagval, l->binv.p);
}

// t = A_l x_l with the halo of x_l refreshed first (several ranks); level 0 goes through the handle's SpMV, whose kernel
// carries the exchange itself over the peer mailboxes
void level_spmv(vfvm_handle* h, Amg& A, size_t i) {
    Level& l = *A.L[i];
    if (A.distributed && i == 0) {
        vfvm_spmv_impl(h, l.x.p, l.t.p, A.fp32);
        return;
    }
    SpmvArgs a = level_args(h, l);
    if (i == 0 && A.fp32) a.offval32 = h->offval32.p;
    if (l.dist) vfvm_spmv_level_halo(h, a, l.halo, l.x.p, l.t.p);
    else if (l.N) vfvm_spmv_level(h, a, l.x.p, l.t.p);
}

void smooth(vfvm_handle* h, Amg& A, size_t i, const double* b, bool first, double* out) {
    cudaStream_t s = h->stream;
    Level& l = *A.L[i];
    if (!first) level_spmv(h, A, i);
    if (!l.N) return;
    NS_SWITCH(h->n, (k_smooth<NS><<<cdiv(l.N, 128), 128, 0, s>>>(l.N, A.omega, l.binv.p, b, first ? nullptr : l.t.p, l.x.p, out)));
    h->launches++;
}

// ---- fused coarse cycle: host side ----------------------------------------------------------------------------------------------
void emit_program(const Amg& A, size_t i, int bsel, std::vector<FuseOp>& P) {
    auto small = [&](size_t lv) { return A.L[lv]->N <= FUSE_SMALL_N ? 1 : 0; };
    auto op = [&](int code, size_t lv, int bs) { P.push_back(FuseOp{code, (int)lv, bs, small(lv)}); };
    if (i + 1 == A.L.size()) {  // coarsest level: a few sweeps
        for (int k = 0; k < A.coarse_sweeps; k++) {
            if (k > 0) op(FOP_SPMV, i, 0);
            op(k == 0 ? FOP_SMOOTH_FIRST : FOP_SMOOTH, i, bsel);
        }
        return;
    }
    op(FOP_SMOOTH_FIRST, i, bsel);
    for (int k = 1; k < A.sweeps; k++) {
        op(FOP_SPMV, i, 0);
        op(FOP_SMOOTH, i, bsel);
    }
    op(FOP_SPMV, i, 0);
    op(FOP_RESTRICT, i, bsel);
    emit_program(A, i + 1, 0, P);
    if ((int)(i + 1) <= A.wdepth && i + 2 < A.L.size()) {
        op(FOP_SPMV, i + 1, 0);
        op(FOP_WRES, i + 1, 0);
        emit_program(A, i + 1, 1, P);
        op(FOP_WADD, i + 1, 0);
    }
    op(FOP_PROLONG, i, 0);
    for (int k = 1; k < A.sweeps; k++) {
        op(FOP_SPMV, i, 0);
        op(FOP_SMOOTH, i, bsel);
    }
    op(FOP_SPMV, i, 0);
    op(FOP_SMOOTH, i, bsel);
}

void build_fused(vfvm_handle* h, Amg& A) {
    A.fuse_valid = false;
    // measured (profiles/r2_*): on one and two B200s the persistent kernel is not faster than the replayed graph of per-level kernels -- every
    // phase still pays its dependent load chain plus a grid barrier -- so it is opt-in (VFVM_AMG_FUSE=1)
    static const bool off = getenv("VFVM_AMG_FUSE") == nullptr || getenv("VFVM_AMG_NO_FUSE") != nullptr;
    if (const char* e = getenv("VFVM_AMG_FUSE_MAX_N")) A.fuse_max_n = std::max(1ll, atoll(e));
    // the fused kernel takes over where launch latency, not bandwidth, bounds a level: the first level with at most fuse_max_n nodes
    A.fuse_level = 0;
    for (size_t i = 1; i < A.L.size(); i++)
        if (A.L[i]->N <= A.fuse_max_n) {
            A.fuse_level = (int)i;
            break;
        }
    if (const char* e = getenv("VFVM_AMG_FUSE_LEVEL")) A.fuse_level = std::max(0, atoi(e));
    if (off || A.fuse_level <= 0 || (size_t)A.fuse_level >= A.L.size()) return;
    if (A.distributed && !h->peer_ok) return;  // NCCL transport: the exchanges are host-enqueued collectives, the levels stay separate kernels
    if (A.repl_level) return;  // replicated coarse levels: the restriction into them ends in an all-gather, the levels stay separate kernels
    const int nn = (int)h->nb_ranks.size();
    std::vector<LevelDev> lv(A.L.size());
    for (size_t i = 0; i < A.L.size(); i++) {
        const Level& l = *A.L[i];
        LevelDev& d = lv[i];
        memset(&d, 0, sizeof(d));
        d.N = l.N;
        d.Nvec = l.Nvec;
        d.nnz_sell = l.nnz_sell;
        d.Nc = l.Nc;
        d.nslices = l.nslices;
        d.sell_ptr = l.sell_ptr;
        d.colidx = l.colidx;
        d.offval = l.offval;
        d.diagval = l.diagval;
        d.binv = l.binv.p;
        d.agg = l.agg.p;
        d.agg_ptr = l.agg_ptr.p;
        d.agg_nodes = l.agg_nodes.p;
        d.x = l.x.p;
        d.t = l.t.p;
        d.b = l.b.p;
        d.b2 = l.b2.p;
        d.x2 = l.x2.p;
        if (A.distributed && i > 0) {
            for (int r = 0; r <= nn; r++) {
                d.send_ptr[r] = l.halo.send_ptr[r];
                d.recv_ptr[r] = l.halo.recv_ptr[r];
            }
            d.send_idx = l.halo.send_idx.p;
        }
    }
    A.lev_dev.upload(lv.data(), lv.size(), h->stream);
    for (int bs = 0; bs < 2; bs++) {
        std::vector<FuseOp> P;
        emit_program(A, (size_t)A.fuse_level, bs, P);
        A.prog_len[bs] = (int)P.size();
        A.prog_dev[bs].upload(P.data(), P.size(), h->stream);
    }
    if (!A.bar.p) {
        A.bar.alloc(2);
        CK(cudaMemsetAsync(A.bar.p, 0, 2 * sizeof(unsigned long long), h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    A.fuse_valid = true;
}

template <int NS, bool DIAGMASK>
void launch_fused_k(vfvm_handle* h, Amg& A, const FuseArgs& fa) {
    auto kern = k_fused_cycle<NS, DIAGMASK>;
    static int occ = 0;
    if (occ == 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FUSE_THREADS, 0));
        if (occ < 1) throw std::string("fused AMG cycle kernel cannot be launched");
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    // one block per SM at most (the blocks meet at barriers: ~1 us among 148 blocks, several among 600); no more blocks than the entry
    // level has slices per warp
    (void)occ;
    const int grid = std::max(1, std::min(nsm, cdiv(A.L[A.fuse_level]->nslices, FUSE_THREADS / 32)));
    kern<<<grid, FUSE_THREADS, 0, h->stream>>>(fa);
    h->launches++;
}

void launch_fused(vfvm_handle* h, Amg& A, int bsel) {
    FuseArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.lev = A.lev_dev.p;
    fa.prog = A.prog_dev[bsel].p;
    fa.nops = A.prog_len[bsel];
    fa.distributed = A.distributed ? 1 : 0;
    fa.omega = A.omega;
    fa.alpha = A.alpha;
    fa.bar_ctr = A.bar.p;
    fa.bar_gen = A.bar.p + 1;
    if (A.distributed) fa.P = vfvm_peer_args_halo(h);
    SpmvArgs sa = vfvm_spmv_args(h);
    memcpy(fa.idxF, sa.idxF, sizeof(fa.idxF));
    memcpy(fa.idxD, sa.idxD, sizeof(fa.idxD));
    bool diagmask = (h->cF == h->n && h->cD == h->n);
    for (int i = 0; i < h->n && diagmask; i++) diagmask = (h->idxF[i * h->n + i] == i && h->idxD[i * h->n + i] == i);
    if (diagmask) {
        NS_SWITCH(h->n, (launch_fused_k<NS, true>(h, A, fa)));
    } else {
        NS_SWITCH(h->n, (launch_fused_k<NS, false>(h, A, fa)));
    }
}

void cycle(vfvm_handle* h, Amg& A, size_t i, const double* b, double* out) {
    cudaStream_t s = h->stream;
    Level& l = *A.L[i];
    if (A.fuse_valid && i >= 1 && (int)i == A.fuse_level && !out && (b == l.b.p || b == l.b2.p)) {  // this level and everything below: one persistent kernel
        launch_fused(h, A, b == l.b.p ? 0 : 1);
        return;
    }
    if (i + 1 == A.L.size()) {  // coarsest level: a few sweeps
        for (int k = 0; k < A.coarse_sweeps; k++) smooth(h, A, i, b, k == 0, k + 1 == A.coarse_sweeps ? out : nullptr);
        return;
    }
    Level& c = *A.L[i + 1];
    // transfer into a replicated level: my aggregates are segment `rank` of its vectors; the right-hand side is completed by one all-gather
    const bool to_repl = A.repl_level && (int)(i + 1) == A.repl_level;
    const int64_t cN = to_repl ? l.Nc : c.N, coff = to_repl ? (int64_t)h->rank * A.repl_S * h->n : 0;
    smooth(h, A, i, b, true, nullptr);
    for (int k = 1; k < A.sweeps; k++) smooth(h, A, i, b, false, nullptr);
    level_spmv(h, A, i);
    if (cN) NS_SWITCH(h->n, (k_restrict<NS><<<cdiv(cN, 128), 128, 0, s>>>(cN, l.agg_ptr.p, l.agg_nodes.p, b, l.t.p, c.b.p + coff)));
    if (to_repl) vfvm_allgather_segments(h, c.b.p, A.repl_S * h->n);
    cycle(h, A, i + 1, c.b.p, nullptr);
    if ((int)(i + 1) <= A.wdepth && i + 2 < A.L.size()) {  // second visit: correct x_c by a cycle on its residual
        const int64_t nd = c.N * h->n;
        level_spmv(h, A, i + 1);
        if (nd) k_w_residual<<<cdiv(nd, 256), 256, 0, s>>>(nd, c.b.p, c.t.p, c.x.p, c.b2.p, c.x2.p);
        cycle(h, A, i + 1, c.b2.p, nullptr);
        if (nd) k_w_add<<<cdiv(nd, 256), 256, 0, s>>>(nd, c.x2.p, c.x.p);
        h->launches += 2;
    }
    if (l.N) NS_SWITCH(h->n, (k_prolong<NS><<<cdiv(l.N, 256), 256, 0, s>>>(l.N, A.alpha, l.agg.p, c.x.p + coff, l.x.p)));
    h->launches += 2;
    for (int k = 1; k < A.sweeps; k++) smooth(h, A, i, b, false, nullptr);
    smooth(h, A, i, b, false, out);
}

}  // namespace

void vfvm_amg_setup(vfvm_handle* h) {
    if (!h->nzfac.p) throw std::string("AMG needs the edge-factor plane of the pattern");
    if (!h->amg) h->amg = new Amg();
    Amg& A = *(Amg*)h->amg;
    A.owner = h;
    {  // tuning knobs (read at every setup so that a probe can sweep them)
        const double theta_old = A.theta;
        if (const char* e = getenv("VFVM_AMG_OMEGA")) A.omega = atof(e);
        if (const char* e = getenv("VFVM_AMG_ALPHA")) A.alpha = atof(e);
        if (const char* e = getenv("VFVM_AMG_THETA")) A.theta = atof(e);
        if (const char* e = getenv("VFVM_AMG_COARSE_SWEEPS")) A.coarse_sweeps = std::max(1, atoi(e));
        if (const char* e = getenv("VFVM_AMG_SWEEPS")) A.sweeps = std::max(1, atoi(e));
        if (const char* e = getenv("VFVM_AMG_WDEPTH")) A.wdepth = std::max(0, atoi(e));
        const int ml_old = A.max_levels;
        if (const char* e = getenv("VFVM_AMG_MAX_LEVELS")) A.max_levels = std::max(1, atoi(e));
        if (A.max_levels != ml_old) A.struct_valid = false;
        if (A.theta != theta_old) A.struct_valid = false;
    }
    if (!A.struct_valid || A.pattern_nnz != h->nnz_sell || A.pattern_N != h->Nown || A.L.empty() || A.L[0]->sell_ptr != h->sell_ptr.p) build_hierarchy(h, A);
    numeric_setup(h, A);
    Amg::Signature sig;
    sig.omega = A.omega;
    sig.alpha = A.alpha;
    sig.sweeps = A.sweeps;
    sig.coarse_sweeps = A.coarse_sweeps;
    sig.wdepth = A.wdepth;
    sig.off0 = h->offval.p;
    sig.diag0 = h->diagval.p;
    sig.off32 = h->offval32.p;
    if (!(sig == A.captured)) {  // a new Jacobian in the same buffers keeps the captured cycles; new options or buffers do not
        A.drop_graphs();
        A.captured = sig;
        A.fuse_valid = false;
    }
    if (!A.fuse_valid) build_fused(h, A);
}

// options of the AMG preconditioner: omega (smoother damping), alpha (weight of the coarse correction), theta (strength threshold),
// sweeps (pre = post smoothing steps), coarse_sweeps, wdepth (levels 1..wdepth are visited twice: W-cycle on the top of the
// hierarchy; measured: halves the CG iterations of the 3D Laplace problem cfg3, 92 -> 65 ms, but costs more than it saves on
// cfg1/2/4/5, hence 0 by default).  NaN keeps the current value.
extern "C" int vfvm_amg_set_options(vfvm_handle* h, const double* opts, int nopts) {
    if (!h || !opts || nopts < 0 || nopts > 6) return VFVM_ERR_ARG;
    if (!h->amg) h->amg = new Amg();
    Amg& A = *(Amg*)h->amg;
    A.owner = h;
    h->graph_epoch++;
    auto have = [&](int k) { return k < nopts && opts[k] == opts[k]; };
    if (have(0)) A.omega = opts[0];
    if (have(1)) A.alpha = opts[1];
    if (have(2) && opts[2] != A.theta) {
        A.theta = opts[2];
        A.struct_valid = false;
    }
    if (have(3)) A.sweeps = std::max(1, (int)opts[3]);
    if (have(4)) A.coarse_sweeps = std::max(1, (int)opts[4]);
    if (have(5)) A.wdepth = std::max(0, (int)opts[5]);
    h->precon_valid = false;
    return VFVM_OK;
}

void vfvm_amg_apply(vfvm_handle* h, const double* in, double* out) {
    Amg& A = *(Amg*)h->amg;
    static const bool no_graph = getenv("VFVM_AMG_NO_GRAPH") != nullptr;
    if (h->in_capture) {  // part of a captured Krylov iteration (linsolve.cu): plain kernels
        cycle(h, A, 0, in, out);
        return;
    }
    if (A.distributed || h->nranks > 1 || no_graph || A.applies++ < 2) {  // the first applications run eagerly (occupancy queries, allocations)
        cycle(h, A, 0, in, out);
        return;
    }
    for (CycleGraph& g : A.graphs)
        if (g.in == in && g.out == out) {
            CK(cudaGraphLaunch(g.exec, h->stream));
            h->launches += g.launches;
            return;
        }
    if (A.graphs.size() >= 8) {  // unexpected number of vector pairs: stay eager
        cycle(h, A, 0, in, out);
        return;
    }
    CycleGraph g;
    g.in = in;
    g.out = out;
    const int64_t before = h->launches;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    try {
        cycle(h, A, 0, in, out);
    } catch (...) {
        cudaStreamEndCapture(h->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    CK(cudaStreamEndCapture(h->stream, &graph));
    g.launches = h->launches - before;
    CK(cudaGraphInstantiate(&g.exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    A.graphs.push_back(g);
    CK(cudaGraphLaunch(g.exec, h->stream));
}

void vfvm_amg_free(vfvm_handle* h) {
    if (h->amg) delete (Amg*)h->amg;
    h->amg = nullptr;
}
