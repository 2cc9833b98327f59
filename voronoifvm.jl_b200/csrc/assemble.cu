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
    const double* __restrict__ Q;    // node-transformed unknowns q(u), n planes of N, for fluxes that are functions of q (flux_node_transform); else = U
    double* __restrict__ F;
    double* __restrict__ offval;
    double* __restrict__ diagval;
    const PhysicsDev* __restrict__ ph;
    int32_t* flags;
    int64_t nnz_sell, Nown, Ntot;
    int slice0, nslices, cF, cD, the_region;  // this launch handles the slices [slice0, nslices)
    double time, tstepinv, lambda;
    signed char idxF[100], idxD[100];
    // masked systems (species not enabled in every cell region): enabled-species bits per cell region, species defined per node
    unsigned short rsmask[VFVM_MAX_CREGIONS];
    const int32_t* __restrict__ node_active;
};

// internal flux id: power-law diffusion with exponent exactly 2 (Example207): u*u, no pow() code in the kernel
#define FLUX_POWDIFF_SQ 100
__host__ __device__ constexpr bool flux_separable(int flux) {
    return flux == VFVM_NONE || flux == VFVM_FLUX_DIFFUSION || flux == VFVM_FLUX_POWDIFF || flux == FLUX_POWDIFF_SQ;
}

// Fluxes of the form g(q(u_K), q(u_L)) with an expensive node map q: q is tabulated once per assembly by k_node_transform
// (2 exp per NODE instead of 4 Dual<6> exp per (row, neighbour) pair) and the neighbour loop gathers q_L instead of u_L.
// Bipolar Scharfetter-Gummel (examples/Example161_BipolarDriftDiffusionCurrent.jl:134-150): q = (n_n, n_p, psi) with
// n_n = exp(z_n (phi_n - psi + E_n)), n_p = exp(z_p (phi_p - psi + E_p)).
__host__ __device__ constexpr bool flux_node_transform(int flux) { return flux == VFVM_FLUX_SG_BIPOLAR; }

template <int FLUX, int NS>
__global__ void k_node_transform(int64_t K0, int64_t K1, int64_t N, const double* __restrict__ U, const PhysicsDev* __restrict__ ph, double* __restrict__ Q) {
    const int64_t K = K0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= K1) return;
    const double* __restrict__ p = ph->params + ph->slot[VFVM_SLOT_FLUX].off;
    if constexpr (FLUX == VFVM_FLUX_SG_BIPOLAR && NS == 3) {
        const double zn = p[3], zp = p[4], En = p[5], Ep = p[6];
        const double phin = U[K * 3], phip = U[K * 3 + 1], psi = U[K * 3 + 2];
        Q[K] = exp(zn * (phin - psi + En));  // one plane per component: the row kernel's gathers are unit-stride across lanes
        Q[N + K] = exp(zp * (phip - psi + Ep));
        Q[2 * N + K] = psi;
    }
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
// MASKED (only with MULTIREG, !SEP): species are enabled per cell region (enable_species!(sys, i, regions)).  An (edge, region)
// or (node, region) item contributes to species i only where region_species[i, region] holds, and to the coupling (i,j) only
// where both hold (assemble_res_jac, src/vfvm_assemblydata.jl:245-270, 350-382); dofs of species that are not defined at a node
// get the identity row F = u - uold, A_ii = 1 (_eval_and_assemble_inactive_species, src/vfvm_system.jl:1012-1026).
template <int NS, int FLUX, bool MULTIREG, bool SEP, bool LIGHT, bool MASKED = false>
__global__ void __launch_bounds__(ASM_THREADS, (SEP && LIGHT && NS == 1) ? 4 : ((!SEP && NS <= 3 && !MASKED) ? 2 : 1)) k_assemble_rows(const AsmArgs a) {
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

    for (int g = a.slice0 + blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
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
            uo_r[i] = (has_storage || MASKED) ? a.UOld[r * NS + i] : 0.0;
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
                // masked: form factor summed over the cell regions of the edge in which species i / both species (i,j) are enabled
                double wsp[MASKED ? NS : 1], wpair[MASKED ? NS * NS : 1];
                if constexpr (MASKED) {
#pragma unroll
                    for (int i = 0; i < NS; i++) wsp[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < NS * NS; i++) wpair[i] = 0.0;
                    const int ed = a.nz_edge[e];
                    if (ed >= 0) {
                        for (int64_t q = a.ef_colptr[ed]; q < a.ef_colptr[ed + 1]; q++) {
                            const unsigned m = a.rsmask[a.ef_region[q] - 1];
                            const double fq = a.ef_fac[q];
#pragma unroll
                            for (int i = 0; i < NS; i++) {
                                if (!((m >> i) & 1u)) continue;
                                wsp[i] += fq;
#pragma unroll
                                for (int jj = 0; jj < NS; jj++)
                                    if ((m >> jj) & 1u) wpair[i * NS + jj] += fq;
                            }
                        }
                    }
                }
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
                            if constexpr (MASKED) {
                                Fr[i] += wsp[i] * f.v;
                                const int p = a.idxF[i * NS + i];
                                if (p >= 0) {
                                    Dr[i * NS + i] += wsp[i] * f.d[0];
                                    a.offval[(int64_t)p * nnz + e] = wsp[i] * f.d[1];
                                }
                                continue;
                            }
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
                            const double si = MASKED ? (pos ? wsp[i] : -wsp[i]) : sfac;
                            Fr[i] += si * f[i].v;
#pragma unroll
                            for (int jj = 0; jj < NS; jj++) {
                                const int p = a.idxF[i * NS + jj];
                                if (p < 0) continue;
                                const double drow = pos ? f[i].d[jj] : f[i].d[NS + jj], dcol = pos ? f[i].d[NS + jj] : f[i].d[jj];
                                nan_seen |= (drow != drow) | (dcol != dcol);
                                const double sij = MASKED ? (pos ? wpair[i * NS + jj] : -wpair[i * NS + jj]) : sfac;
                                Dr[i * NS + jj] += sij * drow;
                                if (first) a.offval[(int64_t)p * nnz + e] = sij * dcol;
                                else a.offval[(int64_t)p * nnz + e] += sij * dcol;
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
                    const unsigned rm = MASKED ? a.rsmask[region - 1] : 0xffffu;
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        if (MASKED && !((rm >> i) & 1u)) continue;
                        Fr[i] += fac * (rea[i].v - srcv[i] + (stor[i].v - ostor[i]) * a.tstepinv);
#pragma unroll
                        for (int jj = 0; jj < NS; jj++) {
                            if (MASKED && !((rm >> jj) & 1u)) continue;
                            const double jv = rea[i].d[jj] + stor[i].d[jj] * a.tstepinv;
                            nan_seen |= (jv != jv);
                            Dr[i * NS + jj] += jv * fac;
                        }
                    }
                }
                if constexpr (MASKED) {  // dofs of species that are not defined at this node: identity row
                    const unsigned act = (unsigned)a.node_active[r];
#pragma unroll
                    for (int i = 0; i < NS; i++)
                        if (!((act >> i) & 1u)) {
                            Fr[i] = u_r[i] - uo_r[i];
#pragma unroll
                            for (int jj = 0; jj < NS; jj++) Dr[i * NS + jj] = (i == jj) ? 1.0 : 0.0;
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
// BT: entries per lane in flight (0 = default for the chunk size); rows of 2D grids have at most 6 neighbours, a batch of 6 covers
// them in one pass without the spills of the 8-wide default
template <int NS, int CH, int FLUX, bool LIGHT, int BT = 0>
__global__ void __launch_bounds__(ASM_THREADS, ((LIGHT && CH == 1) || CH == 2) ? 4 : (CH <= 5 ? 2 : 1)) k_assemble_rows_sep(const AsmArgs a) {
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

    for (int g = a.slice0 + blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
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
        constexpr int BATCH = BT ? BT : (CH == 1 ? 8 : (CH <= 3 ? 4 : 2));
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

// ---- bipolar Scharfetter-Gummel drift-diffusion (cfg4; examples/Example161_BipolarDriftDiffusionCurrent.jl:134-150) ----------
// One edge of g(q_K, q_L), q = (n_n, n_p, psi) tabulated by k_node_transform, with the chain rule through q written out.
// The only dual-number evaluation left is the Bernoulli pair in its single argument x = psi_K - psi_L (B(-x) = x + B(x),
// so bm' = 1 + bp').  POS: the row node is edge.node[1] = K (the larger index); the reference's orientation is kept because
// x + B(x) is not an antisymmetric floating-point expression.  The five flux couplings (n,n) (n,psi) (p,p) (p,psi) (psi,psi)
// are the fixed mask of this flux (pattern.cu), so their planes are plain pointers.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct BipolarCoef {
    double cn, cp, zn, zp, lam2;
};

// B(x) and B'(x) for a batch of independent arguments, branches as src/vfvm_functions.jl:78-90.  The two expensive
// branches are guarded by warp votes and evaluated for the whole batch (instruction-level parallelism across the batch,
// no divergence); the per-lane choice is a select.  |x| >= 0.25: x / expm1(x) with expm1(x) = exp(x) - 1, whose relative
// error there is below 5e-16 (exp(x) - 1 >= 0.22 exp(x)), and d expm1 = exp(x) as ForwardDiff has it.
template <int B>
__device__ __forceinline__ void bernoulli_batch(const double (&x)[B], double (&bp)[B], double (&dbp)[B]) {
    bool small[B], any_small = false, any_big = false;
#pragma unroll
    for (int b = 0; b < B; b++) {
        small[b] = fabs(x[b]) < 0.25;
        any_small |= small[b];
        any_big |= !small[b];
    }
    double hv[B], hd[B], ev[B], ed[B];
#pragma unroll
    for (int b = 0; b < B; b++) hv[b] = hd[b] = ev[b] = ed[b] = 0.0;
    if (__any_sync(0xffffffffu, any_small)) {
#pragma unroll
        for (int b = 0; b < B; b++) {
            Dual<1> X(x[b]);
            X.d[0] = 1.0;
            const Dual<1> Y = bernoulli_horner(X);
            hv[b] = Y.v;
            hd[b] = Y.d[0];
        }
    }
    if (__any_sync(0xffffffffu, any_big)) {
#pragma unroll
        for (int b = 0; b < B; b++) {
            const double e = exp(x[b]), ib = 1.0 / (e - 1.0), y = x[b] * ib;
            ev[b] = y;
            ed[b] = (1.0 - y * e) * ib;
        }
    }
#pragma unroll
    for (int b = 0; b < B; b++) {
        double v = small[b] ? hv[b] : ev[b], d = small[b] ? hd[b] : ed[b];
        if (x[b] < -50.0) {
            v = -x[b];
            d = -1.0;
        }
        if (x[b] > 50.0) v = d = 0.0;
        bp[b] = v;
        dbp[b] = d;
    }
}

template <bool MULTIREG, int BATCH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_assemble_rows_bipolar(const AsmArgs a) {
    constexpr int NS = 3;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const PhysicsDev& ph = *a.ph;
    const double* __restrict__ pf = ph.params + ph.slot[VFVM_SLOT_FLUX].off;
    const int rid = ph.slot[VFVM_SLOT_REACTION].id, sid = ph.slot[VFVM_SLOT_STORAGE].id;
    const double* __restrict__ pr = ph.params + ph.slot[VFVM_SLOT_REACTION].off;
    const double* __restrict__ ps = ph.params + ph.slot[VFVM_SLOT_STORAGE].off;
    const bool has_storage = sid != VFVM_NONE;
    const int64_t nnz = a.nnz_sell;
    const BipolarCoef c = {-pf[3] * pf[1], -pf[4] * pf[2], pf[3], pf[4], pf[0] * pf[0]};
    double* __restrict__ o0 = a.offval + (int64_t)a.idxF[0] * nnz;  // (n,n)
    double* __restrict__ o1 = a.offval + (int64_t)a.idxF[2] * nnz;  // (n,psi)
    double* __restrict__ o2 = a.offval + (int64_t)a.idxF[4] * nnz;  // (p,p)
    double* __restrict__ o3 = a.offval + (int64_t)a.idxF[5] * nnz;  // (p,psi)
    double* __restrict__ o4 = a.offval + (int64_t)a.idxF[8] * nnz;  // (psi,psi)
    bool nan_seen = false;

    for (int g = a.slice0 + blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double Fr[3] = {0.0, 0.0, 0.0}, Df[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        const double* __restrict__ Q0 = a.Q;
        const double* __restrict__ Q1 = a.Q + a.Ntot;
        const double* __restrict__ Q2 = a.Q + 2 * a.Ntot;
        const double nn_r = Q0[r], np_r = Q1[r], ps_r = Q2[r];
        // the node terms' inputs are only needed after the neighbour loop: pull them towards L2 now, without holding registers
        prefetch_l2(a.U + r * 3);
        if (has_storage) prefetch_l2(a.UOld + r * 3);
        if (a.src) prefetch_l2(a.src + r * 3);
        int nq0 = (int)r, nq1 = (int)r + 1;
        if constexpr (MULTIREG) {
            nq0 = (int)a.nf_colptr[r];
            nq1 = (int)a.nf_colptr[r + 1];
        }
        const int gnext = g + nwarps < a.nslices ? g + nwarps : g;  // the slice this warp handles next
        const int base_next = a.sell_ptr[gnext];
        const double zr = c.zn * nn_r, yr = c.zp * np_r;

        // Software pipeline over batches of BATCH entries, three batches in flight: while batch k is computed, the gathers
        // of q_c and the factors of batch k+1 and the indices of batch k+2 are outstanding.  (q, fac, pos) ping-pong
        // between two register sets (A, B), so nothing waits on a load before its batch is due.
        int L[BATCH];
        double fA[BATCH], fB[BATCH], qA[BATCH][3], qB[BATCH][3];
        bool posA[BATCH], posB[BATCH];
        auto load_idx = [&](int jb) {
#pragma unroll
            for (int b = 0; b < BATCH; b++) L[b] = (jb + b < w) ? a.colidx[base + (jb + b) * 32 + lane] : (int)r;  // warp-uniform guard
        };
        auto gather = [&](double (&q)[BATCH][3], double (&f)[BATCH], bool (&pos)[BATCH], int jb) {
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                q[b][0] = Q0[L[b]];
                q[b][1] = Q1[L[b]];
                q[b][2] = Q2[L[b]];
                pos[b] = r > L[b];  // row node is edge.node[1] = K (the larger index)
                f[b] = (jb + b < w) ? a.nzfac[base + (jb + b) * 32 + lane] : 0.0;
            }
        };
        auto compute = [&](const double (&q)[BATCH][3], const double (&f)[BATCH], const bool (&pos)[BATCH], int jb) {
            double x[BATCH], bp[BATCH], dbp[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) x[b] = pos[b] ? ps_r - q[b][2] : q[b][2] - ps_r;  // psi_K - psi_L
            bernoulli_batch<BATCH>(x, bp, dbp);
#pragma unroll
            for (int b = 0; b < BATCH; b++) {
                if (jb + b >= w) break;
                const int e = base + (jb + b) * 32 + lane;
                // Row contribution of edge {r,c} in either orientation (K = larger index; the reference's flux is
                // c_n (bm n_L - bp n_K), c_p (bp p_L - bm p_K) with bp = B(x), bm = B(-x) = x + B(x), x = psi_K - psi_L):
                //   R_n = fac c_n (bc n_c - br n_r),  R_p = fac c_p (br p_c - bc p_r),  (bc, br) = pos ? (bm, bp) : (bp, bm)
                // -- bitwise the reference's value, because negating a difference is exact.
                const double bm = x[b] + bp[b], dbm = 1.0 + dbp[b];
                const double bc = pos[b] ? bm : bp[b], br = pos[b] ? bp[b] : bm;
                const double dc = pos[b] ? dbm : -dbp[b], dr = pos[b] ? dbp[b] : -dbm;  // d/d psi_r of bc, br
                const double nn_c = q[b][0], np_c = q[b][1];
                const double sn = f[b] * c.cn, sp = f[b] * c.cp, sl = f[b] * c.lam2;
                const double zc = c.zn * nn_c, yc = c.zp * np_c;
                const double dn = dc * nn_c - dr * nn_r, dp = dr * np_c - dc * np_r;
                Fr[0] += sn * (bc * nn_c - br * nn_r);
                Fr[1] += sp * (br * np_c - bc * np_r);
                Fr[2] += sl * (ps_r - q[b][2]);
                Df[0] -= sn * (br * zr);
                Df[1] += sn * (dn + br * zr);
                Df[2] -= sp * (bc * yr);
                Df[3] += sp * (dp + bc * yr);
                Df[4] += sl;
                const double v0 = sn * (bc * zc), v1 = -(sn * (dn + bc * zc)), v2 = sp * (br * yc), v3 = -(sp * (dp + br * yc));
                nan_seen |= (v0 != v0) | (v1 != v1) | (v2 != v2) | (v3 != v3);
                o0[e] = v0;
                o1[e] = v1;
                o2[e] = v2;
                o3[e] = v3;
                o4[e] = -sl;
            }
        };
        load_idx(0);
        gather(qA, fA, posA, 0);
        load_idx(BATCH);
        for (int j0 = 0; j0 < w; j0 += 2 * BATCH) {
            gather(qB, fB, posB, j0 + BATCH);
            load_idx(j0 + 2 * BATCH);
            compute(qA, fA, posA, j0);
            gather(qA, fA, posA, j0 + 2 * BATCH);
            load_idx(j0 + 3 * BATCH);
            compute(qB, fB, posB, j0 + BATCH);
        }
        if (gnext != g) {  // head of the next slice's streams -> L2 while the node terms are evaluated
            prefetch_l2(a.colidx + base_next + lane);
            prefetch_l2(a.nzfac + base_next + lane);
            prefetch_l2(a.nzfac + base_next + 32 + lane);
        }
        nan_seen |= (Df[0] != Df[0]) | (Df[1] != Df[1]) | (Df[2] != Df[2]) | (Df[3] != Df[3]);

        // ---------------- node terms (K4) + write-out
        if (valid) {
            double Dr[9] = {Df[0], 0.0, Df[1], 0.0, Df[2], Df[3], 0.0, 0.0, Df[4]};
            typedef Dual<NS> DN;
            DN u[NS];
            double uo[NS], srcv[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                u[i] = DN(a.U[r * NS + i]);
                u[i].d[i] = 1.0;
                uo[i] = has_storage ? a.UOld[r * NS + i] : 0.0;
                srcv[i] = a.src ? a.src[r * NS + i] : 0.0;
            }
            for (int q = nq0; q < nq1; q++) {
                const double fac = a.nf_fac[q];
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
    if (nan_seen) atomicOr(a.flags, 1);
}

// ---- edgereaction(f,u,edge,data), src/vfvm_assembly.jl:202-239 ------------------------------------------------------------------------
// A separate pass over the same SELL-32 rows after the row kernel (a feature path: the hot kernels stay untouched).  The reference adds
// fac f to BOTH end nodes' residuals and (K,K) += J_K, (L,K) -= J_K, (K,L) -= J_L, (L,L) += J_L with K = edge.node[1]; seen from row r
// with neighbour c: F_r += fac f, diag_r += fac df/du_r, offdiag(r,c) -= fac df/du_c.
template <int NS>
__global__ void k_edgereaction_rows(const AsmArgs a, const double* __restrict__ coord, int dim) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const PhysicsDev& ph = *a.ph;
    const int eid = ph.slot[VFVM_SLOT_EDGEREACTION].id;
    const double* __restrict__ pe = ph.params + ph.slot[VFVM_SLOT_EDGEREACTION].off;
    const int64_t nnz = a.nnz_sell;
    bool nan_seen = false;
    for (int g = a.slice0 + blockIdx.x * wpb + (threadIdx.x >> 5); g < a.nslices; g += nwarps) {
        const int64_t rraw = (int64_t)g * 32 + lane;
        const bool valid = rraw < a.Nown;
        const int64_t r = valid ? rraw : a.Nown - 1;
        const int base = a.sell_ptr[g];
        const int w = (a.sell_ptr[g + 1] - base) >> 5;
        double u_r[NS], Fr[NS], Dr[NS * NS], xr[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u_r[i] = a.U[r * NS + i];
            Fr[i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < NS * NS; i++) Dr[i] = 0.0;
        for (int d = 0; d < dim; d++) xr[d] = coord[r * dim + d];
        for (int j = 0; j < w; j++) {
            const int64_t e = (int64_t)base + (int64_t)j * 32 + lane;
            const int L = a.colidx[e];
            const double fac = a.nzfac[e];
            if (!valid || L == r) continue;  // padding
            double h2 = 0.0;
            for (int d = 0; d < dim; d++) {
                const double dx = xr[d] - coord[(int64_t)L * dim + d];
                h2 += dx * dx;
            }
            const bool pos = r > L;  // row node is edge.node[1]
            typedef Dual<2 * NS> D;
            D x[NS], y[NS], f[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                const double uc = a.U[(int64_t)L * NS + i];
                x[i] = D(pos ? u_r[i] : uc);
                x[i].d[i] = 1.0;
                y[i] = D(pos ? uc : u_r[i]);
                y[i].d[NS + i] = 1.0;
                f[i] = D(0.0);
            }
            eval_edgereaction<NS>(eid, pe, f, x, y, sqrt(h2), dim);
#pragma unroll
            for (int i = 0; i < NS; i++) {
                Fr[i] += fac * f[i].v;
#pragma unroll
                for (int jj = 0; jj < NS; jj++) {
                    const int p = a.idxF[i * NS + jj];
                    if (p < 0) continue;
                    const double drow = pos ? f[i].d[jj] : f[i].d[NS + jj], dcol = pos ? f[i].d[NS + jj] : f[i].d[jj];
                    nan_seen |= (drow != drow) | (dcol != dcol);
                    Dr[i * NS + jj] += fac * drow;
                    a.offval[(int64_t)p * nnz + e] -= fac * dcol;
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                a.F[r * NS + i] += Fr[i];
#pragma unroll
                for (int jj = 0; jj < NS; jj++) {
                    const int pD = a.idxD[i * NS + jj];
                    if (pD >= 0 && a.idxF[i * NS + jj] >= 0) a.diagval[(int64_t)pD * a.Nown + r] += Dr[i * NS + jj];
                }
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
    const double* __restrict__ UOld;
    double tstepinv;
    double* __restrict__ F;
    double* __restrict__ diagval;
    const PhysicsDev* __restrict__ ph;
    int32_t* flags;
    int64_t b0, nbnodes, Nown;  // this launch handles the boundary nodes [b0, nbnodes)
    int dim;
    double time, lambda;
    signed char idxD[100];
    const int32_t* __restrict__ node_active;  // masked systems: species defined at the node (isnodespecies, src/vfvm_assemblydata.jl:286-292), else null
};

template <int NS>
__global__ void k_assemble_bnodes(const BNodeArgs a) {
    const int64_t b = a.b0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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
    const unsigned act = a.node_active ? (unsigned)a.node_active[K] : 0xffffu;
    for (int q = a.bn_ptr[b]; q < a.bn_ptr[b + 1]; q++) {
        const int ibf = a.bn_bface[q];
        const int region = a.bfaceregions[ibf];
        const double fac = a.bfnf[(int64_t)ibf * a.dim + a.bn_local[q]];
        double Dirichlet = 1.0e30;  // src/vfvm_geometryitems.jl:296
        if (ph.has_legacy_bc) {     // src/vfvm_assembly.jl:355-387
            Dirichlet = 1.0e30 / fac;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                if (!((act >> i) & 1u)) continue;
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
            if (!((act >> i) & 1u)) continue;
            Fk[i] += fac * res[i].v;
#pragma unroll
            for (int j = 0; j < NS; j++) {
                if (!((act >> j) & 1u)) continue;
                const double jv = res[i].d[j];
                nan_seen |= (jv != jv);
                const int pD = a.idxD[i * NS + j];
                if (pD >= 0 && jv != 0.0) a.diagval[(int64_t)pD * a.Nown + K] += jv * fac;
            }
        }
        if (ph.slot[VFVM_SLOT_BSTORAGE].id != VFVM_NONE) {  // src/vfvm_assembly.jl:409-439
            DN st[NS];
            double uo[NS], sto[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                st[i] = DN(0.0);
                sto[i] = 0.0;
                uo[i] = a.UOld[(int64_t)K * NS + i];
            }
            eval_bstorage<NS>(ph, st, u, region);
            eval_bstorage<NS>(ph, sto, uo, region);
#pragma unroll
            for (int i = 0; i < NS; i++) {
                if (!((act >> i) & 1u)) continue;
                Fk[i] += fac * (st[i].v - sto[i]) * a.tstepinv;
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    if (!((act >> j) & 1u)) continue;
                    const double jv = st[i].d[j];
                    nan_seen |= (jv != jv);
                    const int pD = a.idxD[i * NS + j];
                    if (pD >= 0 && jv != 0.0) a.diagval[(int64_t)pD * a.Nown + K] += jv * (fac * a.tstepinv);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++) a.F[(int64_t)K * NS + i] = Fk[i];
    if (nan_seen) atomicOr(a.flags, 1);
}

// _initialize_inactive_dof! src/vfvm_system.jl:1030-1042: dofs of species that are not defined at a node are zero
__global__ void k_zero_inactive(int64_t Nown, int ns, const int32_t* __restrict__ node_active, double* __restrict__ U) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= Nown) return;
    const unsigned act = (unsigned)node_active[K];
    for (int i = 0; i < ns; i++)
        if (!((act >> i) & 1u)) U[K * ns + i] = 0.0;
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
            if (a.node_active && !(((unsigned)a.node_active[K] >> i) & 1u)) continue;
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
static void launch_slices(vfvm_handle* h, Kern kern, int& plan, const AsmArgs& a, int max_threads = ASM_THREADS) {
    if (plan == 0) {
        int best_t = 0, best_w = 0, best_b = 0;
        for (int t = max_threads; t >= 64; t /= 2) {
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
    const int grid = std::max(1, std::min(cdiv(a.nslices - a.slice0, threads / 32), nsm * occ * h->grid_pct / 100));
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
    const int gx = std::max(1, std::min(cdiv(a.nslices - a.slice0, threads / 32), std::max(1, nsm * occ * h->grid_pct / 100 / nchunks)));
    kern<<<dim3(gx, nchunks), threads, 0, h->stream>>>(a);
    h->launches++;
}

// ---- host dispatch -------------------------------------------------------------------------------------------
// the separable fast path applies when every coupling mask is exactly the species diagonal
static bool fast_path_ok(const vfvm_handle* h) {
    if (h->masked) return false;
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
                    if constexpr (NS == 1) {
                        if (h->group_maxnnz <= 6 && !getenv("VFVM_SEP_BATCH8")) {  // 2D grids: one batch of 6 per row
                            static int plan6 = 0;
                            launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX_POWDIFF_SQ, true, 6>, plan6, a, NS / CH);
                            return;
                        }
                    }
                    launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX_POWDIFF_SQ, true>, plan[2], a, NS / CH);
                    return;
                }
            }
            if constexpr (NS == 10 && FLUX == VFVM_FLUX_DIFFUSION) {
                if (light && getenv("VFVM_SEP_CH2")) {  // tuning probe: 5 chunks of 2 species, 4 blocks per SM
                    static int plan2 = 0;
                    launch_slices_sep(h, k_assemble_rows_sep<NS, 2, FLUX, true>, plan2, a, NS / 2);
                    return;
                }
            }
            if (light && FLUX == VFVM_FLUX_DIFFUSION) launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX, true>, plan[1], a, NS / CH);
            else launch_slices_sep(h, k_assemble_rows_sep<NS, CH, FLUX, false>, plan[0], a, NS / CH);
            return;
        }
    }
    if constexpr (FLUX == VFVM_FLUX_SG_BIPOLAR && NS == 3) {
        if (a.Q != a.U) {
            // 128-thread blocks x 3 per SM (168 registers, no spills in the neighbour loop) measured 5-8 % faster than 256 x 2
            static int ps = 0, pb = 0;
            // Round 2 re-measured the occupancy / batch trade on cfg4 at 193^3 (profiles/r2_bipolar_variants.txt): <2,128,4> (128 registers,
            // 244 B spills, 16 warps/SM) 1.89 ms, <1,128,5> (96 registers, 20 warps/SM) 2.15 ms, <3,128,3> 1.81 ms against 1.79 ms for this one
            if (h->single_region) launch_slices(h, k_assemble_rows_bipolar<false, 2, 128, 3>, ps, a, 128);
            else launch_slices(h, k_assemble_rows_bipolar<true, 2, 128, 3>, pb, a, 128);
            return;
        }
    }
    if constexpr (flux_supported(FLUX, NS)) {
        static int occ0 = 0, occ1 = 0, occm = 0;
        if (h->masked) launch_slices(h, k_assemble_rows<NS, FLUX, true, false, false, true>, occm, a);
        else if (h->single_region) launch_slices(h, k_assemble_rows<NS, FLUX, false, false, false>, occ0, a);
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
        case VFVM_FLUX_MIXTURE: launch_rows<NS, VFVM_FLUX_MIXTURE>(h, a); break;
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

static void fill_asm_args(vfvm_handle* h, AsmArgs& a, double time, double tstepinv, double lambda) {
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
    a.Q = a.U;
    a.F = h->vec[VFVM_VEC_RESIDUAL].p;
    a.offval = h->offval.p;
    a.diagval = h->diagval.p;
    a.ph = h->phys_dev.p;
    a.flags = h->flags.p + 1;  // word 1: NaN seen during assembly (word 0 belongs to the linear solver)
    a.nnz_sell = h->nnz_sell;
    a.Nown = h->Nown;
    a.Ntot = h->N;
    a.slice0 = 0;
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
    for (int r = 0; r < VFVM_MAX_CREGIONS; r++) {
        unsigned short m = 0;
        for (int i = 0; i < h->n && r < h->ncellregions; i++)
            if (h->region_species[(size_t)r * h->n + i]) m |= (unsigned short)(1u << i);
        a.rsmask[r] = m;
    }
    a.node_active = h->node_active.p;
    if (flux_node_transform(h->phys.slot[VFVM_SLOT_FLUX].id) && !h->masked && !getenv("VFVM_GENERIC_DUAL_FLUX")) {  // env: parity probe of the generic Dual<2n> path
        h->node_q.alloc((size_t)h->n * h->N);
        a.Q = h->node_q.p;
    }
}

static void fill_bnode_args(vfvm_handle* h, BNodeArgs& b, const AsmArgs& a, double time, double lambda) {
    b.bn_node = h->bn_node.p;
    b.bn_ptr = h->bn_ptr.p;
    b.bn_bface = h->bn_bface.p;
    b.bn_local = h->bn_local.p;
    b.bfaceregions = h->bfaceregions.p;
    b.bfnf = h->bfacenodefac.p;
    b.U = h->vec[VFVM_VEC_SOLUTION].p;
    b.UOld = a.UOld;
    b.tstepinv = a.tstepinv;
    b.F = h->vec[VFVM_VEC_RESIDUAL].p;
    b.diagval = h->diagval.p;
    b.ph = a.ph;
    b.flags = h->flags.p + 1;
    b.b0 = 0;
    b.nbnodes = h->nbnodes;
    b.Nown = h->Nown;
    b.dim = h->dim;
    b.time = time;
    b.lambda = lambda;
    memcpy(b.idxD, a.idxD, sizeof(b.idxD));
    b.node_active = h->masked ? h->node_active.p : nullptr;
}

// q(u) of the nodes [K0, K1) (only for flux_node_transform fluxes)
static void launch_node_transform(vfvm_handle* h, const AsmArgs& a, int64_t K0, int64_t K1) {
    if (a.Q == a.U || K1 <= K0) return;
    if (h->n == 3) k_node_transform<VFVM_FLUX_SG_BIPOLAR, 3><<<cdiv(K1 - K0, 256), 256, 0, h->stream>>>(K0, K1, h->N, a.U, a.ph, h->node_q.p);
    h->launches++;
}
static void launch_rows_range(vfvm_handle* h, AsmArgs a, int s0, int s1) {
    if (s1 <= s0) return;
    a.slice0 = s0;
    a.nslices = s1;
    NS_DISPATCH(h->n, (launch_rows_ns<NS>(h, a)));
    if (h->phys.slot[VFVM_SLOT_EDGEREACTION].id != VFVM_NONE) {
        if (h->masked) throw std::string("edgereaction with species enabled per region has no device instantiation");
        const int grid = std::max(1, std::min(cdiv(s1 - s0, ASM_WARPS), 148 * 4));
        NS_DISPATCH(h->n, (k_edgereaction_rows<NS><<<grid, ASM_THREADS, 0, h->stream>>>(a, h->coord.p, h->dim)));
        h->launches++;
    }
}
static void launch_bnodes_range(vfvm_handle* h, BNodeArgs b, int64_t b0, int64_t b1) {
    if (b1 <= b0) return;
    b.b0 = b0;
    b.nbnodes = b1;
    NS_DISPATCH(h->n, (k_assemble_bnodes<NS><<<cdiv(b1 - b0, 128), 128, 0, h->stream>>>(b)));
    h->launches++;
}

int vfvm_assemble_impl(vfvm_handle* h, double time, double tstep, double lambda, bool async) {
    cudaStream_t s = h->stream;
    const double tstepinv = 1.0 / tstep;  // src/vfvm_assembly.jl:554 (1/Inf == 0)
    if (tstepinv != 0.0) h->seen_transient = true;
    AsmArgs a;
    fill_asm_args(h, a, time, tstepinv, lambda);
    if (!async) CK(cudaMemsetAsync(h->flags.p + 1, 0, sizeof(int32_t), s));  // async: the flag stays sticky until vfvm_sync
    CK(cudaEventRecord(h->ev0, s));
    launch_node_transform(h, a, 0, h->N);
    launch_rows_range(h, a, 0, h->ngroups);
    CK(cudaEventRecord(h->ev1, s));
    if (h->nbnodes) {
        BNodeArgs b;
        fill_bnode_args(h, b, a, time, lambda);
        launch_bnodes_range(h, b, 0, h->nbnodes);
    }
    CK(cudaEventRecord(h->ev2, s));
    // the preconditioner of the previous Jacobian stays valid for `reuse_precs` solves (factorize_every_newtonstep = false,
    // src/vfvm_solver.jl:99): it is rebuilt whenever vfvm_linsolve is called without reuse, and invalidated by pattern / option changes
    h->asm_pending = true;
    if (async) return VFVM_OK;
    return vfvm_assemble_finish(h);
}

// ---- evaluate_residual_and_jacobian with HOST vectors, pipelined (src/vfvm_solver.jl:224-243) ---------------------------
// The rows are cut into chunks of whole slices.  U travels host -> device chunk by chunk on a copy stream; the rows of chunk c
// are assembled as soon as the last piece of U their columns reach has arrived (for a banded numbering: piece c+1), and
// the residual of chunk c returns to the host on a second copy stream while later chunks are still being uploaded and
// assembled.  PCIe is full duplex, so the call approaches max(H2D, D2H) instead of H2D + assembly + D2H.
// per slice: the largest OWNED column its rows reach (>= the rows themselves), and whether any column is a halo node (>= Nown)
__global__ void k_slice_colmax(int nslices, int64_t Nown, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ colidx, int32_t* __restrict__ out,
                               int32_t* __restrict__ out_halo) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= nslices) return;
    int m = (int)min((int64_t)g * 32 + lane, Nown - 1);  // the row itself
    int hl = 0;
    for (int e = sell_ptr[g] + lane; e < sell_ptr[g + 1]; e += 32) {
        const int c = colidx[e];
        if (c >= Nown) hl = 1;
        else m = max(m, c);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        hl |= __shfl_xor_sync(0xffffffffu, hl, o);
    }
    if (lane == 0) {
        out[g] = m;
        out_halo[g] = hl;
    }
}

static int pipe_chunks(const vfvm_handle* h) {  // pieces of >= 4 MB keep the copies at full PCIe rate
    const int64_t bytes = (int64_t)sizeof(double) * h->n * h->N;
    const char* env = getenv("VFVM_PIPE_CHUNKS");
    const int K = env ? atoi(env) : (int)std::min<int64_t>(16, bytes / (7 << 19));  // measured on cfg3 (57.5 MB each way): 16-24 chunks of >= 3.5 MB are best
    return std::max(1, std::min(K, h->ngroups));
}

static void build_pipe_plan(vfvm_handle* h) {
    PipePlan& P = h->pipe;
    const int K = pipe_chunks(h);
    P.K = K;
    P.slice_begin.assign(K + 1, 0);
    for (int c = 0; c <= K; c++) P.slice_begin[c] = (int)((int64_t)h->ngroups * c / K);
    DevBuf<int32_t> cm, ch;
    cm.alloc(h->ngroups);
    ch.alloc(h->ngroups);
    k_slice_colmax<<<cdiv(h->ngroups, 8), 256, 0, h->stream>>>(h->ngroups, h->Nown, h->sell_ptr.p, h->colidx.p, cm.p, ch.p);
    h->launches++;
    std::vector<int32_t> colmax = cm.to_host(h->stream), colhalo = ch.to_host(h->stream);
    P.piece_hi.assign(K, 0);
    P.needs_halo.assign(K, 0);
    P.bn_begin.assign(K + 1, 0);
    for (int c = 0; c < K; c++) {
        int32_t m = 0;
        for (int g = P.slice_begin[c]; g < P.slice_begin[c + 1]; g++) {
            m = std::max(m, colmax[g]);
            if (colhalo[g]) P.needs_halo[c] = 1;
        }
        int p = c;
        while (p + 1 < K && (int64_t)P.slice_begin[p + 1] * 32 <= m) p++;
        P.piece_hi[c] = p;
        P.bn_begin[c + 1] = std::lower_bound(h->bn_node_host.begin(), h->bn_node_host.end(), (int32_t)std::min<int64_t>((int64_t)P.slice_begin[c + 1] * 32, h->Nown)) - h->bn_node_host.begin();
    }
    P.bn_begin[K] = h->nbnodes;
    if (!h->stream_in) {
        CK(cudaStreamCreateWithFlags(&h->stream_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
    }
    while ((int)h->pipe_ev.size() < 2 * K + 2) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pipe_ev.push_back(e);
    }
    P.valid = true;
}

bool vfvm_pipeline_applies(vfvm_handle* h) {
    if (getenv("VFVM_NO_PIPELINE")) return false;
    const char* env = getenv("VFVM_PIPE_MIN_BYTES");  // below this vector size one copy each way is as fast
    return (int64_t)sizeof(double) * h->n * h->N >= (env ? atoll(env) : (8ll << 20));
}

int vfvm_eval_res_jac_pipelined(vfvm_handle* h, const double* U, const double* UOld, double* F, double time, double tstep, double lambda) {
    if (!h->pipe.valid || h->pipe.K != pipe_chunks(h)) build_pipe_plan(h);
    const PipePlan& P = h->pipe;
    cudaStream_t s = h->stream, sin = h->stream_in, sout = h->stream_out;
    const double tstepinv = 1.0 / tstep;
    if (tstepinv != 0.0) h->seen_transient = true;
    AsmArgs a;
    fill_asm_args(h, a, time, tstepinv, lambda);
    BNodeArgs bn;
    fill_bnode_args(h, bn, a, time, lambda);
    const int n = h->n, K = P.K;
    // UOld == NULL means UOld = U: the kernels read U in its place; the resident OLDSOL vector is refreshed by one device copy
    // after the last chunk (a device copy per piece on the copy-in stream would compete with the copy engines)
    const bool have_old = UOld && UOld != U;
    if (!have_old) a.UOld = a.U;
    bn.UOld = a.UOld;
    // owned pieces [node_begin(p), node_begin(p + 1)); with several ranks the halo nodes [Nown, N) form one more piece that travels FIRST: the rows
    // next to a partition boundary sit in the first and the last chunk, and a halo at the end of the queue would serialise the whole call
    auto node_begin = [&](int c) { return c >= K ? h->Nown : std::min<int64_t>((int64_t)P.slice_begin[c] * 32, h->Nown); };
    const int64_t nhalo = h->N - h->Nown;
    CK(cudaMemsetAsync(h->flags.p + 1, 0, sizeof(int32_t), s));
    CK(cudaEventRecord(h->ev0, s));
    CK(cudaEventRecord(h->pipe_ev[2 * K], s));  // the copy-in stream must not overwrite U before earlier work on the main stream is done
    CK(cudaStreamWaitEvent(sin, h->pipe_ev[2 * K], 0));
    double* dU = h->vec[VFVM_VEC_SOLUTION].p;
    double* dO = h->vec[VFVM_VEC_OLDSOL].p;
    double* dF = h->vec[VFVM_VEC_RESIDUAL].p;
    if (nhalo > 0) {
        const int64_t o = h->Nown * n, cnt = nhalo * n;
        CK(cudaMemcpyAsync(dU + o, U + o, cnt * sizeof(double), cudaMemcpyHostToDevice, sin));
        if (have_old) CK(cudaMemcpyAsync(dO + o, UOld + o, cnt * sizeof(double), cudaMemcpyHostToDevice, sin));
        CK(cudaEventRecord(h->pipe_ev[2 * K + 1], sin));
    }
    for (int p = 0; p < K; p++) {
        const int64_t o = node_begin(p) * n, cnt = (node_begin(p + 1) - node_begin(p)) * n;
        if (cnt > 0) {
            CK(cudaMemcpyAsync(dU + o, U + o, cnt * sizeof(double), cudaMemcpyHostToDevice, sin));
            if (have_old) CK(cudaMemcpyAsync(dO + o, UOld + o, cnt * sizeof(double), cudaMemcpyHostToDevice, sin));
        }
        CK(cudaEventRecord(h->pipe_ev[p], sin));
    }
    // A row kernel on the full persistent grid saturates HBM and starves the copy engines (measured: the upload stalls for
    // exactly the kernels' run time).  On 30 % of the grid the chunk kernels still keep up with PCIe and the copies run at
    // full rate beside them (cfg3: 2.51 ms unpipelined, 1.76 ms pipelined on the full grid, 1.5 ms on 30 %).
    const char* gp = getenv("VFVM_PIPE_GRID_PCT");
    struct GridShare {  // restored on every exit path, also when a launch throws
        vfvm_handle* h;
        ~GridShare() { h->grid_pct = 100; }
    } grid_share{h};
    h->grid_pct = gp ? std::max(1, std::min(100, atoi(gp))) : 30;
    int arrived = -1;  // last piece the main stream has waited for
    bool halo_arrived = nhalo == 0;
    for (int c = 0; c < K; c++) {
        if (!halo_arrived && P.needs_halo[c]) {
            CK(cudaStreamWaitEvent(s, h->pipe_ev[2 * K + 1], 0));
            launch_node_transform(h, a, h->Nown, h->N);
            halo_arrived = true;
        }
        if (P.piece_hi[c] > arrived) {
            CK(cudaStreamWaitEvent(s, h->pipe_ev[P.piece_hi[c]], 0));
            launch_node_transform(h, a, node_begin(arrived + 1), node_begin(P.piece_hi[c] + 1));
            arrived = P.piece_hi[c];
        }
        launch_rows_range(h, a, P.slice_begin[c], P.slice_begin[c + 1]);
        launch_bnodes_range(h, bn, P.bn_begin[c], P.bn_begin[c + 1]);
        CK(cudaEventRecord(h->pipe_ev[K + c], s));
        CK(cudaStreamWaitEvent(sout, h->pipe_ev[K + c], 0));
        const int64_t r0 = (int64_t)P.slice_begin[c] * 32, r1 = std::min<int64_t>((int64_t)P.slice_begin[c + 1] * 32, h->Nown);
        CK(cudaMemcpyAsync(F + r0 * n, dF + r0 * n, (r1 - r0) * n * sizeof(double), cudaMemcpyDeviceToHost, sout));
    }
    h->grid_pct = 100;
    CK(cudaEventRecord(h->ev1, s));
    CK(cudaEventRecord(h->ev2, s));
    if (!have_old) CK(cudaMemcpyAsync(dO, dU, sizeof(double) * n * h->N, cudaMemcpyDeviceToDevice, s));
    h->asm_pending = true;
    static const bool trace = getenv("VFVM_PIPE_TRACE") != nullptr;  // diagnostic: when each stream finished, relative to the start
    cudaEvent_t t_in = nullptr, t_out = nullptr;
    if (trace) {
        CK(cudaEventCreate(&t_in));
        CK(cudaEventCreate(&t_out));
        CK(cudaEventRecord(t_in, sin));
        CK(cudaEventRecord(t_out, sout));
    }
    CK(cudaStreamSynchronize(sout));
    if (trace) {
        CK(cudaStreamSynchronize(s));
        float a_in = 0, a_cmp = 0, a_out = 0;
        cudaEventElapsedTime(&a_in, h->ev0, t_in);
        cudaEventElapsedTime(&a_cmp, h->ev0, h->ev2);
        cudaEventElapsedTime(&a_out, h->ev0, t_out);
        fprintf(stderr, "[vfvm pipe] K=%d  copy-in done %.3f ms, last chunk assembled %.3f ms, copy-out done %.3f ms\n", K, a_in, a_cmp, a_out);
        cudaEventDestroy(t_in);
        cudaEventDestroy(t_out);
    }
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

// masked systems: dofs of species that are not defined at a node are kept at exactly zero (a Krylov solve of their identity rows
// may leave rounding noise)
void vfvm_zero_inactive(vfvm_handle* h, double* vec) {
    if (!h->masked) return;
    k_zero_inactive<<<cdiv(h->Nown, 256), 256, 0, h->stream>>>(h->Nown, h->n, h->node_active.p, vec);
    h->launches++;
}

int vfvm_init_dirichlet_impl(vfvm_handle* h, double time, double lambda) {
    vfvm_zero_inactive(h, h->vec[VFVM_VEC_SOLUTION].p);
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
    b.node_active = h->masked ? h->node_active.p : nullptr;
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

// ---- parity probe of the device inplace_linsolve! (test/test040_inplacelu.jl) ----------------------------------------------------------
template <int N>
__global__ void k_probe_linsolve(int nsys, int pivoting, const double* __restrict__ A, const double* __restrict__ b, double* __restrict__ x) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    double M[N * N], r[N];
    for (int i = 0; i < N * N; i++) M[i] = A[(size_t)s * N * N + i];
    for (int i = 0; i < N; i++) r[i] = b[(size_t)s * N + i];
    if (pivoting) inplace_linsolve_piv<N>(M, r);
    else inplace_linsolve_nopiv<N>(M, r);
    for (int i = 0; i < N; i++) x[(size_t)s * N + i] = r[i];
}

extern "C" int vfvm_probe_inplace_linsolve(vfvm_handle* h, int n, int nsys, int pivoting, const double* A, const double* b, double* x) {
    if (!h || n < 1 || n > 10 || nsys < 0 || !A || !b || !x) return VFVM_ERR_ARG;
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        DevBuf<double> dA, db, dx;
        dA.upload(A, (size_t)nsys * n * n, h->stream);
        db.upload(b, (size_t)nsys * n, h->stream);
        dx.alloc((size_t)std::max(1, nsys) * n);
        const int grid = std::max(1, cdiv(nsys, 64));
        switch (n) {
            case 1: k_probe_linsolve<1><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 2: k_probe_linsolve<2><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 3: k_probe_linsolve<3><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 4: k_probe_linsolve<4><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 5: k_probe_linsolve<5><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 6: k_probe_linsolve<6><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 7: k_probe_linsolve<7><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 8: k_probe_linsolve<8><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            case 9: k_probe_linsolve<9><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
            default: k_probe_linsolve<10><<<grid, 64, 0, h->stream>>>(nsys, pivoting, dA.p, db.p, dx.p); break;
        }
        h->launches++;
        if (nsys) CK(cudaMemcpyAsync(x, dx.p, sizeof(double) * nsys * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
    })
    return VFVM_OK;
}
