// K4 + K5 + K6: residual + Jacobian assembly (eval_and_assemble, src/vfvm_assembly.jl:520-643).
//
// B200 design (not the reference's edge loop): ONE fused kernel streams the off-diagonal blocks of the Jacobian once, in
// the SELL-32 order the pattern stores them in.  A warp owns a slice of 32 consecutive node rows, one lane per row:
// in step j every lane handles the j-th neighbour L of its row K -- (colidx, form factor) loads and the Jacobian-block
// store are perfectly coalesced, the gather of u_L is coalesced whenever neighbouring nodes are numbered consecutively
// (tensor grids, bandwidth-reducing orderings).  The lane evaluates the flux of edge {K,L} in forward-mode duals with the
// reference's orientation (edge.node[1] = larger node), writes the off-diagonal block (K,L) and accumulates the residual
// and diagonal-block contributions of row K in registers, in column order (deterministic; no atomics, no shared memory,
// no barriers, no memset of the matrix, no edge->nnz scatter map).  The node terms (source, reaction, storage: K4) are
// added by the same lane before F and the diagonal block are written once.  Every edge is evaluated from both ends
// (2x flops on a bandwidth-bound kernel) in exchange for write-once coalesced traffic.  The boundary-node kernel (K6)
// runs afterwards, one thread per boundary node over its (bface, local node) items in the reference's loop order.
//
//   assemble_nodes   src/vfvm_assembly.jl:38-126    -> node part of k_assemble_rows
//   assemble_edges   src/vfvm_assembly.jl:128-200   -> neighbour loop of k_assemble_rows
//   assemble_bnodes  src/vfvm_assembly.jl:318-407   -> k_assemble_bnodes
//   _addnz NaN check src/vfvm_assembly.jl:21-24     -> flags[0]
#include <algorithm>

#include "physics.cuh"
#include "vfvm_internal.h"

#define ASM_THREADS 256
#define ASM_WARPS (ASM_THREADS / 32)

struct AsmArgs {
    const int32_t* __restrict__ sell_ptr;
    const int32_t* __restrict__ colidx;
    const double* __restrict__ nzfac;
    const int32_t* __restrict__ nz_edge;
    const int64_t* __restrict__ ef_colptr;
    const int32_t* __restrict__ ef_region;
    const double* __restrict__ ef_fac;
    const int64_t* __restrict__ nf_colptr;
    const int32_t* __restrict__ nf_region;
    const double* __restrict__ nf_fac;
    const double* __restrict__ U;
    const double* __restrict__ UOld;
    const double* __restrict__ src;  // cached source term n x N (u-independent callback), or null
    double* __restrict__ F;
    double* __restrict__ offval;
    double* __restrict__ diagval;
    const PhysicsDev* __restrict__ ph;
    int32_t* flags;
    int64_t nnz_sell, Nown;
    int nslices, cF, cD, the_region;
    double time, tstepinv, lambda;
    signed char idxF[100], idxD[100];
};

// internal flux id: power-law diffusion with exponent exactly 2 (Example207): u*u, no pow() code in the kernel
#define FLUX_POWDIFF_SQ 100
__host__ __device__ constexpr bool flux_separable(int flux) {
    return flux == VFVM_NONE || flux == VFVM_FLUX_DIFFUSION || flux == VFVM_FLUX_POWDIFF || flux == FLUX_POWDIFF_SQ;
}

// species-separable, antisymmetric fluxes f_i = D_i (g(u_i,K) - g(u_i,L)): 2 partials instead of 2n.  fac*f and its
// derivatives are bitwise the same for either edge orientation, so no orientation handling is needed for them.
template <int FLUX>
__device__ __forceinline__ Dual<2> eval_flux_sep(double Di, double m, const Dual<2>& a, const Dual<2>& b) {
    if constexpr (FLUX == VFVM_FLUX_DIFFUSION) return Di * (a - b);
    else if constexpr (FLUX == VFVM_FLUX_POWDIFF) return Di * (dpowr(a, m) - dpowr(b, m));
    else if constexpr (FLUX == FLUX_POWDIFF_SQ) return Di * (a * a - b * b);
    else return Dual<2>(0.0);
}
// LIGHT: node physics without transcendental code paths (power reactions with exponent 1 or 2, affine, linear storage):
// keeps pow()/exp() out of the streaming kernels of cfg1/2/3/5 (register pressure)
template <bool LIGHT, class T>
__device__ __forceinline__ T reaction_sep(int id, const double* __restrict__ p, int i, int ns, const T& u) {
    if constexpr (LIGHT) {
        if (id == VFVM_REACTION_POW) return (p[ns + i] == 2.0) ? p[i] * (u * u) : p[i] * u;
        if (id == VFVM_REACTION_AFFINE) return p[1] + p[0] * u;
        return T(0.0);
    }
    switch (id) {
        case VFVM_REACTION_POW: return p[i] * dpowr(u, p[ns + i]);
        case VFVM_REACTION_SINH: return p[i] * (dexp(u) - dexp(-u));
        case VFVM_REACTION_AFFINE: return p[1] + p[0] * u;  // ns == 1 only
        default: return T(0.0);
    }
}
template <bool LIGHT, class T>
__device__ __forceinline__ T storage_sep(int id, const double* __restrict__ p, int i, int ns, const T& u) {
    if constexpr (LIGHT) return id == VFVM_STORAGE_LINEAR ? p[i] * u : T(0.0);
    switch (id) {
        case VFVM_STORAGE_LINEAR: return p[i] * u;
        case VFVM_STORAGE_POW: return dpowr(p[i] + u, 1.0 / p[ns + i]);
        default: return T(0.0);
    }
}

// SEP: every coupling mask is exactly the species diagonal (plane i <-> (i,i)) and flux / reaction / storage are species-
// separable: the fast path of cfg1/2/3/5.  Otherwise the general path with Dual<2 NS> and the runtime plane tables.
template <int NS, int FLUX, bool MULTIREG, bool SEP, bool LIGHT>
__global__ void __launch_bounds__(ASM_THREADS, (SEP && LIGHT && NS == 1) ? 4 : ((!SEP && NS <= 3) ? 2 : 1)) k_assemble_rows(const AsmArgs a) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const PhysicsDev& ph = *a.ph;
    const double* __restrict__ pf = ph.params + ph.slot[VFVM_SLOT_FLUX].off;
    const int rid = ph.slot[VFVM_SLOT_REACTION].id, sid = ph.slot[VFVM_SLOT_STORAGE].id;
    const double* __restrict__ pr = ph.params + ph.slot[VFVM_SLOT_REACTION].off;
    const double* __restrict__ ps = ph.params + ph.slot[VFVM_SLOT_STORAGE].off;
    const bool has_storage = sid != VFVM_NONE;
    const int64_t nnz = a.nnz_sell;
    bool nan_seen = false;
    double Dcoef[NS], mexp = 0.0;  // separable fluxes: coefficients live in registers
#pragma unroll
    for (int i = 0; i < NS; i++) Dcoef[i] = flux_separable(FLUX) && FLUX != VFVM_NONE ? pf[i] : 0.0;
    if constexpr (FLUX == VFVM_FLUX_POWDIFF) mexp = pf[NS];

    for (int g = blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double u_r[NS], Fr[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u_r[i] = a.U[r * NS + i];
            Fr[i] = 0.0;
        }
        // node data of this lane's row, loaded now so that the latency overlaps the neighbour loop
        double nfac0 = 0.0, uo_r[NS], src_r[NS];
        if constexpr (!MULTIREG) nfac0 = a.nf_fac[r];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            uo_r[i] = has_storage ? a.UOld[r * NS + i] : 0.0;
            src_r[i] = a.src ? a.src[r * NS + i] : 0.0;
        }
        constexpr int ND = SEP ? NS : NS * NS;
        double Dr[ND];
#pragma unroll
        for (int i = 0; i < ND; i++) Dr[i] = 0.0;

        // ---------------- neighbour loop (K5): BATCH entries per lane in flight (index/factor loads, then gathers, then math)
        constexpr int BATCH = SEP ? (NS == 1 ? 8 : (NS <= 3 ? 4 : 2)) : (NS == 1 ? 4 : (NS == 2 ? 2 : 1));
        for (int j0 = 0; j0 < w; j0 += BATCH) {
            int Lc[BATCH];
            double fc[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                const bool ok = j0 + b < w;  // warp-uniform
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
                Lc[b] = ok ? a.colidx[e] : (int)r;
                fc[b] = ok ? a.nzfac[e] : 0.0;
            }
            double ucb[BATCH][NS];
#pragma unroll
            for (int b = 0; b < BATCH; b++)
#pragma unroll
                for (int i = 0; i < NS; i++) ucb[b][i] = a.U[(int64_t)Lc[b] * NS + i];
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (j0 + b >= w) break;
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
                const int L = Lc[b];
                const double* uc = ucb[b];
                const double fac = fc[b];
                {
                    constexpr bool first = true;
                    if constexpr (flux_separable(FLUX)) {
#pragma unroll
                        for (int i = 0; i < NS; i++) {
                            Dual<2> x(u_r[i]), y(uc[i]);
                            x.d[0] = 1.0;
                            y.d[1] = 1.0;
                            const Dual<2> f = eval_flux_sep<FLUX>(Dcoef[i], mexp, x, y);
                            nan_seen |= (f.d[0] != f.d[0]) | (f.d[1] != f.d[1]);
                            Fr[i] += fac * f.v;
                            if constexpr (SEP) {
                                Dr[i] += fac * f.d[0];
                                if (first) a.offval[(int64_t)i * nnz + e] = fac * f.d[1];
                                else a.offval[(int64_t)i * nnz + e] += fac * f.d[1];
                            } else {
                                const int p = a.idxF[i * NS + i];
                                if (p >= 0) {
                                    Dr[i * NS + i] += fac * f.d[0];
                                    if (first) a.offval[(int64_t)p * nnz + e] = fac * f.d[1];
                                    else a.offval[(int64_t)p * nnz + e] += fac * f.d[1];
                                }
                            }
                        }
                    } else {
                        const bool pos = r > L;  // row node is edge.node[1] (the larger index): flux(u_row, u_col), sign +
                        const double sfac = pos ? fac : -fac;
                        typedef Dual<2 * NS> D;
                        D x[NS], y[NS], f[NS];
#pragma unroll
                        for (int i = 0; i < NS; i++) {
                            x[i] = D(pos ? u_r[i] : uc[i]);
                            x[i].d[i] = 1.0;
                            y[i] = D(pos ? uc[i] : u_r[i]);
                            y[i].d[NS + i] = 1.0;
                            f[i] = D(0.0);
                        }
                        eval_flux<FLUX, NS>(pf, f, x, y);
#pragma unroll
                        for (int i = 0; i < NS; i++) {
                            Fr[i] += sfac * f[i].v;
#pragma unroll
                            for (int jj = 0; jj < NS; jj++) {
                                const int p = a.idxF[i * NS + jj];
                                if (p < 0) continue;
                                const double drow = pos ? f[i].d[jj] : f[i].d[NS + jj], dcol = pos ? f[i].d[NS + jj] : f[i].d[jj];
                                nan_seen |= (drow != drow) | (dcol != dcol);
                                Dr[i * NS + jj] += sfac * drow;
                                if (first) a.offval[(int64_t)p * nnz + e] = sfac * dcol;
                                else a.offval[(int64_t)p * nnz + e] += sfac * dcol;
                            }
                        }
                    }
                }
            }
        }

        // ---------------- node terms (K4) + write-out
        if (valid) {
            if constexpr (SEP) {
                const double nfac = nfac0;
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    Dual<1> u(u_r[i]);
                    u.d[0] = 1.0;
                    const Dual<1> rea = reaction_sep<LIGHT>(rid, pr, i, NS, u);
                    Dual<1> stor(0.0);
                    double ostor = 0.0;
                    if (has_storage) {
                        stor = storage_sep<LIGHT>(sid, ps, i, NS, u);
                        ostor = storage_sep<LIGHT>(sid, ps, i, NS, uo_r[i]);
                    }
                    const double srcv = src_r[i];
                    const double jv = rea.d[0] + stor.d[0] * a.tstepinv;
                    nan_seen |= (jv != jv);
                    a.F[r * NS + i] = Fr[i] + nfac * (rea.v - srcv + (stor.v - ostor) * a.tstepinv);
                    a.diagval[(int64_t)i * a.Nown + r] = Dr[i] + jv * nfac;
                }
            } else {
                int64_t q0 = r, q1 = r + 1;
                if constexpr (MULTIREG) {
                    q0 = a.nf_colptr[r];
                    q1 = a.nf_colptr[r + 1];
                }
                typedef Dual<NS> DN;
                DN u[NS];
                const double* uo = uo_r;
                const double* srcv = src_r;
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    u[i] = DN(u_r[i]);
                    u[i].d[i] = 1.0;
                }
                for (int64_t q = q0; q < q1; q++) {
                    const double fac = MULTIREG ? a.nf_fac[q] : nfac0;
                    const int region = MULTIREG ? a.nf_region[q] : a.the_region;
                    double ostor[NS];
                    DN rea[NS], stor[NS];
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        ostor[i] = 0.0;
                        rea[i] = DN(0.0);
                        stor[i] = DN(0.0);
                    }
                    eval_reaction<NS>(rid, pr, rea, u, region);
                    if (has_storage) {
                        eval_storage<NS>(sid, ps, stor, u);
                        eval_storage<NS>(sid, ps, ostor, uo);
                    }
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        Fr[i] += fac * (rea[i].v - srcv[i] + (stor[i].v - ostor[i]) * a.tstepinv);
#pragma unroll
                        for (int jj = 0; jj < NS; jj++) {
                            const double jv = rea[i].d[jj] + stor[i].d[jj] * a.tstepinv;
                            nan_seen |= (jv != jv);
                            Dr[i * NS + jj] += jv * fac;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    a.F[r * NS + i] = Fr[i];
#pragma unroll
                    for (int jj = 0; jj < NS; jj++) {
                        const int pD = a.idxD[i * NS + jj];
                        if (pD >= 0) a.diagval[(int64_t)pD * a.Nown + r] = Dr[i * NS + jj];
                    }
                }
            }
        }
    }
    if (nan_seen) atomicOr(a.flags, 1);
}

// Separable fast path (every coupling mask = species diagonal; cfg1/2/3/5).  Species are independent scalar problems on the
// same graph, so they are processed in chunks of CH species (blockIdx.y = chunk): registers stay low for many-species
// systems (cfg5: 10 species -> 2 chunks of 5) at the price of re-reading the 12 B/entry index+factor stream per chunk.
template <int NS, int CH, int FLUX, bool LIGHT>
__global__ void __launch_bounds__(ASM_THREADS, (LIGHT && CH == 1) ? 4 : (CH <= 5 ? 2 : 1)) k_assemble_rows_sep(const AsmArgs a) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const int c0 = blockIdx.y * CH;
    const PhysicsDev& ph = *a.ph;
    const double* __restrict__ pf = ph.params + ph.slot[VFVM_SLOT_FLUX].off;
    const int rid = ph.slot[VFVM_SLOT_REACTION].id, sid = ph.slot[VFVM_SLOT_STORAGE].id;
    const double* __restrict__ pr = ph.params + ph.slot[VFVM_SLOT_REACTION].off;
    const double* __restrict__ ps = ph.params + ph.slot[VFVM_SLOT_STORAGE].off;
    const bool has_storage = sid != VFVM_NONE;
    const int64_t nnz = a.nnz_sell;
    bool nan_seen = false;
    double Dcoef[CH], mexp = 0.0;
#pragma unroll
    for (int i = 0; i < CH; i++) Dcoef[i] = pf[c0 + i];
    if constexpr (FLUX == VFVM_FLUX_POWDIFF) mexp = pf[NS];

    for (int g = blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double u_r[CH], Fr[CH], Dr[CH], uo_r[CH], src_r[CH];
        const double nfac = a.nf_fac[r];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            u_r[i] = a.U[r * NS + c0 + i];
            uo_r[i] = has_storage ? a.UOld[r * NS + c0 + i] : 0.0;
            src_r[i] = a.src ? a.src[r * NS + c0 + i] : 0.0;
            Fr[i] = 0.0;
            Dr[i] = 0.0;
        }
        constexpr int BATCH = CH == 1 ? 8 : (CH <= 3 ? 4 : 2);
        for (int j0 = 0; j0 < w; j0 += BATCH) {
            int Lc[BATCH];
            double fc[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                const bool ok = j0 + b < w;  // warp-uniform
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
                Lc[b] = ok ? a.colidx[e] : (int)r;
                fc[b] = ok ? a.nzfac[e] : 0.0;
            }
            double ucb[BATCH][CH];
#pragma unroll
            for (int b = 0; b < BATCH; b++)
#pragma unroll
                for (int i = 0; i < CH; i++) ucb[b][i] = a.U[(int64_t)Lc[b] * NS + c0 + i];
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (j0 + b >= w) break;
                const int64_t e = (int64_t)base + (int64_t)(j0 + b) * 32 + lane;
#pragma unroll
                for (int i = 0; i < CH; i++) {
                    Dual<2> x(u_r[i]), y(ucb[b][i]);
                    x.d[0] = 1.0;
                    y.d[1] = 1.0;
                    const Dual<2> f = eval_flux_sep<FLUX>(Dcoef[i], mexp, x, y);
                    nan_seen |= (f.d[0] != f.d[0]) | (f.d[1] != f.d[1]);
                    Fr[i] += fc[b] * f.v;
                    Dr[i] += fc[b] * f.d[0];
                    a.offval[(int64_t)(c0 + i) * nnz + e] = fc[b] * f.d[1];
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                Dual<1> u(u_r[i]);
                u.d[0] = 1.0;
                const Dual<1> rea = reaction_sep<LIGHT>(rid, pr, c0 + i, NS, u);
                Dual<1> stor(0.0);
                double ostor = 0.0;
                if (has_storage) {
                    stor = storage_sep<LIGHT>(sid, ps, c0 + i, NS, u);
                    ostor = storage_sep<LIGHT>(sid, ps, c0 + i, NS, uo_r[i]);
                }
                const double jv = rea.d[0] + stor.d[0] * a.tstepinv;
                nan_seen |= (jv != jv);
                a.F[r * NS + c0 + i] = Fr[i] + nfac * (rea.v - src_r[i] + (stor.v - ostor) * a.tstepinv);
                a.diagval[(int64_t)(c0 + i) * a.Nown + r] = Dr[i] + jv * nfac;
            }
        }
    }
    if (nan_seen) atomicOr(a.flags, 1);
}

// tabulates the (u-independent) source callback once per physics change: src[i,K] = source(f, node)[i]
template <int NS>
__global__ void k_source_cache(int64_t N, int dim, const double* __restrict__ coord, const PhysicsDev* __restrict__ ph, double* __restrict__ out) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= N) return;
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) s[i] = 0.0;
    eval_source<NS>(ph->slot[VFVM_SLOT_SOURCE].id, ph->params + ph->slot[VFVM_SLOT_SOURCE].off, s, coord + K * dim, dim, ph->nodal_source, K);
#pragma unroll
    for (int i = 0; i < NS; i++) out[K * NS + i] = s[i];
}

// ---- K6: boundary nodes ---------------------------------------------------------------------------------
struct BNodeArgs {
    const int32_t* __restrict__ bn_node;
    const int32_t* __restrict__ bn_ptr;
    const int32_t* __restrict__ bn_bface;
    const int32_t* __restrict__ bn_local;
    const int32_t* __restrict__ bfaceregions;
    const double* __restrict__ bfnf;
    double* U;  // read (assembly) or written (init_dirichlet)
    double* __restrict__ F;
    double* __restrict__ diagval;
    const PhysicsDev* __restrict__ ph;
    int32_t* flags;
    int64_t nbnodes, Nown;
    int dim;
    double time, lambda;
    signed char idxD[100];
};

template <int NS>
__global__ void k_assemble_bnodes(const BNodeArgs a) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nbnodes) return;
    const PhysicsDev& ph = *a.ph;
    const int K = a.bn_node[b];
    typedef Dual<NS> DN;
    DN u[NS];
    double Fk[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) {
        u[i] = DN(a.U[(int64_t)K * NS + i]);
        u[i].d[i] = 1.0;
        Fk[i] = a.F[(int64_t)K * NS + i];
    }
    bool nan_seen = false;
    for (int q = a.bn_ptr[b]; q < a.bn_ptr[b + 1]; q++) {
        const int ibf = a.bn_bface[q];
        const int region = a.bfaceregions[ibf];
        const double fac = a.bfnf[(int64_t)ibf * a.dim + a.bn_local[q]];
        double Dirichlet = 1.0e30;  // src/vfvm_geometryitems.jl:296
        if (ph.has_legacy_bc) {     // src/vfvm_assembly.jl:355-387
            Dirichlet = 1.0e30 / fac;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                const double bf = ph.bfactors[(region - 1) * NS + i], bv = ph.bvalues[(region - 1) * NS + i];
                const int pD = a.idxD[i * NS + i];
                if (bf == 1.0e30) {
                    Fk[i] += bf * (u[i].v - bv);
                    a.diagval[(int64_t)pD * a.Nown + K] += bf;
                } else {
                    Fk[i] += fac * (bf * u[i].v - bv);
                    if (bf != 0.0) a.diagval[(int64_t)pD * a.Nown + K] += bf * fac;
                }
            }
        }
        DN res[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) res[i] = DN(0.0);
        eval_breaction<NS>(ph, res, u, region, a.time, Dirichlet, (double*)nullptr);
#pragma unroll
        for (int i = 0; i < NS; i++) {  // src/vfvm_assembly.jl:399-401
            Fk[i] += fac * res[i].v;
#pragma unroll
            for (int j = 0; j < NS; j++) {
                const double jv = res[i].d[j];
                nan_seen |= (jv != jv);
                const int pD = a.idxD[i * NS + j];
                if (pD >= 0 && jv != 0.0) a.diagval[(int64_t)pD * a.Nown + K] += jv * fac;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++) a.F[(int64_t)K * NS + i] = Fk[i];
    if (nan_seen) atomicOr(a.flags, 1);
}

// _initialize_dirichlet! src/vfvm_system.jl:947-1003: later items overwrite earlier ones (loop order = bface order)
template <int NS>
__global__ void k_init_dirichlet(const BNodeArgs a) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nbnodes) return;
    const PhysicsDev& ph = *a.ph;
    const int K = a.bn_node[b];
    double u[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) u[i] = a.U[(int64_t)K * NS + i];
    for (int q = a.bn_ptr[b]; q < a.bn_ptr[b + 1]; q++) {
        const int region = a.bfaceregions[a.bn_bface[q]];
        double dv[NS], y[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            dv[i] = INFINITY;
            y[i] = 0.0;
        }
        eval_breaction<NS>(ph, y, u, region, a.time, 1.0e30, dv);
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (!isinf(dv[i])) u[i] = dv[i];
            if (ph.has_legacy_bc) {
                const double bf = ph.bfactors[(region - 1) * NS + i];
                if (fabs(bf - 1.0e30) <= 1.4901161193847656e-8 * fmax(fabs(bf), 1.0e30)) u[i] = ph.bvalues[(region - 1) * NS + i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++) a.U[(int64_t)K * NS + i] = u[i];
}

// persistent grid: one wave of blocks (SMs x resident blocks), warps stride over the slices.  The block size is the one
// that keeps the most warps resident for the kernel's register footprint (heavy dual-number kernels prefer small blocks).
template <class Kern>
static void launch_slices(vfvm_handle* h, Kern kern, int& plan, const AsmArgs& a) {
    if (plan == 0) {
        int best_t = 0, best_w = 0, best_b = 0;
        for (int t = ASM_THREADS; t >= 64; t /= 2) {
            int b = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, t, 0));
            if (b * t > best_w) {
                best_w = b * t;
                best_t = t;
                best_b = b;
            }
        }
        if (best_t == 0) throw std::string("assembly kernel cannot be launched (registers)");
        plan = best_t * 1024 + best_b;
    }
    const int threads = plan / 1024, occ = plan % 1024;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::max(1, std::min(cdiv(a.nslices, threads / 32), nsm * occ));
    kern<<<grid, threads, 0, h->stream>>>(a);
    h->launches++;
}

template <class Kern>
static void launch_slices_sep(vfvm_handle* h, Kern kern, int& plan, const AsmArgs& a, int nchunks) {
    if (plan == 0) {
        int best_t = 0, best_w = 0, best_b = 0;
        for (int t = ASM_THREADS; t >= 64; t /= 2) {
            int b = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, t, 0));
            if (b * t > best_w) {
                best_w = b * t;
                best_t = t;
                best_b = b;
            }
        }
        if (best_t == 0) throw std::string("assembly kernel cannot be launched (registers)");
        plan = best_t * 1024 + best_b;
    }
    const int threads = plan / 1024, occ = plan % 1024;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int gx = std::max(1, std::min(cdiv(a.nslices, threads / 32), std::max(1, nsm * occ / nchunks)));
    kern<<<dim3(gx, nchunks), threads, 0, h->stream>>>(a);
    h->launches++;
}

// ---- host dispatch -------------------------------------------------------------------------------------------
// the separable fast path applies when every coupling mask is exactly the species diagonal
static bool fast_path_ok(const vfvm_handle* h) {
    const int n = h->n, fid = h->phys.slot[VFVM_SLOT_FLUX].id, rid = h->phys.slot[VFVM_SLOT_REACTION].id, sid = h->phys.slot[VFVM_SLOT_STORAGE].id;
    if (!h->single_region || !(fid == VFVM_FLUX_DIFFUSION || fid == VFVM_FLUX_POWDIFF)) return false;
    if (!(rid == VFVM_NONE || rid == VFVM_REACTION_POW || rid == VFVM_REACTION_SINH || (rid == VFVM_REACTION_AFFINE && n == 1))) return false;
    if (!(sid == VFVM_NONE || sid == VFVM_STORAGE_LINEAR || sid == VFVM_STORAGE_POW)) return false;
    if (h->cF != n || h->cD != n) return false;
    for (int i = 0; i < n; i++)
        if (h->idxF[i * n + i] != i || h->idxD[i * n + i] != i) return false;
    return getenv("VFVM_NO_FAST_PATH") == nullptr;
}

// node physics without transcendental functions?
static bool light_node_physics(const vfvm_handle* h) {
    const PhysicsDev& ph = h->phys;
    const int n = h->n, rid = ph.slot[VFVM_SLOT_REACTION].id, sid = ph.slot[VFVM_SLOT_STORAGE].id;
    if (!(sid == VFVM_NONE || sid == VFVM_STORAGE_LINEAR)) return false;
    if (rid == VFVM_NONE || (rid == VFVM_REACTION_AFFINE && n == 1)) return true;
    if (rid == VFVM_REACTION_POW) {
        const double* p = ph.params + ph.slot[VFVM_SLOT_REACTION].off;
        for (int i = 0; i < n; i++)
            if (!(p[n + i] == 1.0 || p[n + i] == 2.0)) return false;
        return true;
    }
    return false;
}

template <int NS, int FLUX>
static void launch_rows(vfvm_handle* h, const AsmArgs& a) {
    if constexpr (FLUX == VFVM_FLUX_DIFFUSION || FLUX == VFVM_FLUX_POWDIFF) {
        if (fast_path_ok(h)) {
            constexpr int CH = NS > 5 ? 5 : NS;
            static_assert(NS % CH == 0, "species chunking");
            static int plan[3] = {0, 0, 0};
            const bool light = light_node_physics(h);
            if constexpr (NS == 1 && FLUX == VFVM_FLUX_DIFFUSION) {
                if (light && !getenv("VFVM_SEP_CHUNKED")) {  // measured on cfg3: the unchunked variant is 10 % faster for one species
                    static int planx = 0;
                    launch_slices(h, k_assemble_rows<NS, FLUX, false, true, true>, planx, a);
                    return;
                }
            }
            if constexpr (FLUX == VFVM_FLUX_POWDIFF) {
                const double m = h->phys.params[h->phys.slot[VFVM_SLOT_FLUX].off + NS];
                if (m == 2.0 && light) {
                    launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX_POWDIFF_SQ, true>, plan[2], a, NS / CH);
                    return;
                }
            }
            if (light && FLUX == VFVM_FLUX_DIFFUSION) launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX, true>, plan[1], a, NS / CH);
            else launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX, false>, plan[0], a, NS / CH);
            return;
        }
    }
    if constexpr (flux_supported(FLUX, NS)) {
        static int occ0 = 0, occ1 = 0;
        if (h->single_region) launch_slices(h, k_assemble_rows<NS, FLUX, false, false, false>, occ0, a);
        else launch_slices(h, k_assemble_rows<NS, FLUX, true, false, false>, occ1, a);
    } else {
        throw std::string("flux id ") + std::to_string(FLUX) + " has no device instantiation for " + std::to_string(NS) + " species";
    }
}

template <int NS>
static void launch_rows_ns(vfvm_handle* h, const AsmArgs& a) {
    switch (h->phys.slot[VFVM_SLOT_FLUX].id) {
        case VFVM_NONE: launch_rows<NS, VFVM_NONE>(h, a); break;
        case VFVM_FLUX_DIFFUSION: launch_rows<NS, VFVM_FLUX_DIFFUSION>(h, a); break;
        case VFVM_FLUX_POWDIFF: launch_rows<NS, VFVM_FLUX_POWDIFF>(h, a); break;
        case VFVM_FLUX_CROSSDIFF2: launch_rows<NS, VFVM_FLUX_CROSSDIFF2>(h, a); break;
        case VFVM_FLUX_SG_UNIPOLAR: launch_rows<NS, VFVM_FLUX_SG_UNIPOLAR>(h, a); break;
        case VFVM_FLUX_SEDAN: launch_rows<NS, VFVM_FLUX_SEDAN>(h, a); break;
        case VFVM_FLUX_SG_BIPOLAR: launch_rows<NS, VFVM_FLUX_SG_BIPOLAR>(h, a); break;
        default: throw std::string("unregistered flux id");
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

// (re)tabulate the source callback; called when the physics block changed
void vfvm_source_cache(vfvm_handle* h) {
    if (h->phys.slot[VFVM_SLOT_SOURCE].id == VFVM_NONE) {
        h->src_cache.release();
        return;
    }
    h->src_cache.alloc((size_t)h->n * h->N);
    NS_DISPATCH(h->n, (k_source_cache<NS><<<cdiv(h->N, 256), 256, 0, h->stream>>>(h->N, h->dim, h->coord.p, h->phys_dev.p, h->src_cache.p)));
    h->launches++;
}

int vfvm_assemble_impl(vfvm_handle* h, double time, double tstep, double lambda, bool async) {
    cudaStream_t s = h->stream;
    const double tstepinv = 1.0 / tstep;  // src/vfvm_assembly.jl:554 (1/Inf == 0)
    if (tstepinv != 0.0) h->seen_transient = true;
    AsmArgs a;
    a.sell_ptr = h->sell_ptr.p;
    a.colidx = h->colidx.p;
    a.nzfac = h->nzfac.p;
    a.nz_edge = h->nz_edge.p;
    a.ef_colptr = h->ef_colptr.p;
    a.ef_region = h->ef_region.p;
    a.ef_fac = h->ef_fac.p;
    a.nf_colptr = h->nf_colptr.p;
    a.nf_region = h->nf_region.p;
    a.nf_fac = h->nf_fac.p;
    a.U = h->vec[VFVM_VEC_SOLUTION].p;
    a.UOld = h->vec[VFVM_VEC_OLDSOL].p;
    a.src = h->src_cache.p;
    a.F = h->vec[VFVM_VEC_RESIDUAL].p;
    a.offval = h->offval.p;
    a.diagval = h->diagval.p;
    a.ph = h->phys_dev.p;
    a.flags = h->flags.p + 1;  // word 1: NaN seen during assembly (word 0 belongs to the linear solver)
    a.nnz_sell = h->nnz_sell;
    a.Nown = h->Nown;
    a.nslices = h->ngroups;
    a.cF = h->cF;
    a.cD = h->cD;
    a.the_region = h->the_region;
    a.time = time;
    a.tstepinv = tstepinv;
    a.lambda = lambda;
    for (int b = 0; b < 100; b++) {
        a.idxF[b] = (signed char)(b < h->n * h->n ? h->idxF[b] : -1);
        a.idxD[b] = (signed char)(b < h->n * h->n ? h->idxD[b] : -1);
    }
    if (!async) CK(cudaMemsetAsync(h->flags.p + 1, 0, sizeof(int32_t), s));  // async: the flag stays sticky until vfvm_sync
    CK(cudaEventRecord(h->ev0, s));
    NS_DISPATCH(h->n, (launch_rows_ns<NS>(h, a)));
    CK(cudaEventRecord(h->ev1, s));
    if (h->nbnodes) {
        BNodeArgs b;
        b.bn_node = h->bn_node.p;
        b.bn_ptr = h->bn_ptr.p;
        b.bn_bface = h->bn_bface.p;
        b.bn_local = h->bn_local.p;
        b.bfaceregions = h->bfaceregions.p;
        b.bfnf = h->bfacenodefac.p;
        b.U = h->vec[VFVM_VEC_SOLUTION].p;
        b.F = h->vec[VFVM_VEC_RESIDUAL].p;
        b.diagval = h->diagval.p;
        b.ph = a.ph;
        b.flags = h->flags.p + 1;
        b.nbnodes = h->nbnodes;
        b.Nown = h->Nown;
        b.dim = h->dim;
        b.time = time;
        b.lambda = lambda;
        memcpy(b.idxD, a.idxD, sizeof(b.idxD));
        NS_DISPATCH(h->n, (k_assemble_bnodes<NS><<<cdiv(h->nbnodes, 128), 128, 0, s>>>(b)));
        h->launches++;
    }
    CK(cudaEventRecord(h->ev2, s));
    h->precon_valid = false;
    h->asm_pending = true;
    if (async) return VFVM_OK;
    return vfvm_assemble_finish(h);
}

// waits for the stream, reads the timing events of the last assembly and the (sticky) NaN flag
int vfvm_assemble_finish(vfvm_handle* h) {
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(h->flags_host + 2, h->flags.p + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (!h->asm_pending) return VFVM_OK;
    h->asm_pending = false;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev2));
    h->times[VFVM_TIME_ASSEMBLE] = ms;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->times[VFVM_TIME_EDGE_KERNEL] = ms;
    if (h->flags_host[2] & 1) {
        CK(cudaMemsetAsync(h->flags.p + 1, 0, sizeof(int32_t), s));
        return vfvm_fail(h, VFVM_ERR_NAN, "trying to assemble NaN");
    }
    return VFVM_OK;
}

int vfvm_init_dirichlet_impl(vfvm_handle* h, double time, double lambda) {
    if (!h->nbnodes) return VFVM_OK;
    BNodeArgs b;
    memset(&b, 0, sizeof(b));
    b.bn_node = h->bn_node.p;
    b.bn_ptr = h->bn_ptr.p;
    b.bn_bface = h->bn_bface.p;
    b.bn_local = h->bn_local.p;
    b.bfaceregions = h->bfaceregions.p;
    b.bfnf = h->bfacenodefac.p;
    b.U = h->vec[VFVM_VEC_SOLUTION].p;
    b.ph = h->phys_dev.p;
    b.nbnodes = h->nbnodes;
    b.Nown = h->Nown;
    b.dim = h->dim;
    b.time = time;
    b.lambda = lambda;
    NS_DISPATCH(h->n, (k_init_dirichlet<NS><<<cdiv(h->nbnodes, 128), 128, 0, h->stream>>>(b)));
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return VFVM_OK;
}

// ---- parity probe of the device Bernoulli function ------------------------------------------------------------------
__global__ void k_probe_bernoulli(int n, const double* __restrict__ x, double* __restrict__ bp, double* __restrict__ bm, double* __restrict__ dbp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Dual<1> xx(x[i]), p, m;
    xx.d[0] = 1.0;
    fbernoulli_pm(xx, p, m);
    bp[i] = p.v;
    bm[i] = m.v;
    dbp[i] = p.d[0];
}

extern "C" int vfvm_probe_bernoulli(vfvm_handle* h, int n, const double* x, double* bp, double* bm, double* dbp) {
    if (!h || n < 0) return VFVM_ERR_ARG;
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        DevBuf<double> dx, d1, d2, d3;
        dx.upload(x, n, h->stream);
        d1.alloc(n);
        d2.alloc(n);
        d3.alloc(n);
        if (n) k_probe_bernoulli<<<cdiv(n, 256), 256, 0, h->stream>>>(n, dx.p, d1.p, d2.p, d3.p);
        h->launches++;
        d1.download(bp, h->stream);
        d2.download(bm, h->stream);
        d3.download(dbp, h->stream);
        CK(cudaGetLastError());
    })
    return VFVM_OK;
}
