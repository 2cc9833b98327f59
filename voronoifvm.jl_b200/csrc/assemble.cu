// K4 + K5 + K6: residual + Jacobian assembly (eval_and_assemble, src/vfvm_assembly.jl:520-643).
//
// B200 design (not the reference's edge loop): ONE fused row-tile kernel streams the off-diagonal block CSR
// once.  A warp owns a group of consecutive node rows; one lane per off-diagonal block (K,L) evaluates the flux of
// the edge {K,L} in forward-mode duals with the reference's orientation (edge.node[1] = larger node), writes the
// off-diagonal Jacobian block with a fully coalesced store (no edge->nnz scatter, no atomics, no memset of the
// matrix) and leaves its residual / diagonal-block contribution in the warp's shared memory; one lane per row then reduces
// the row segment in fixed column order (deterministic), adds the node terms (source, reaction, storage: K4 fused)
// and writes F and the diagonal block once.  Every edge is evaluated from both ends (2x flops on a bandwidth-bound
// kernel) in exchange for write-once coalesced traffic.  The boundary-node kernel (K6) runs afterwards, one thread
// per boundary node over its (bface, local node) items in the reference's loop order.
//
//   assemble_nodes   src/vfvm_assembly.jl:38-126    -> row phase of k_assemble_rows
//   assemble_edges   src/vfvm_assembly.jl:128-200   -> block phase of k_assemble_rows
//   assemble_bnodes  src/vfvm_assembly.jl:318-407   -> k_assemble_bnodes
//   _addnz NaN check src/vfvm_assembly.jl:21-24     -> flags[0]
#include "physics.cuh"
#include "vfvm_internal.h"

#define ASM_THREADS 256
#define ASM_WARPS (ASM_THREADS / 32)

struct AsmArgs {
    const int32_t* __restrict__ rowptr;
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
    int64_t nnz_off, Nown;
    int ngroups, group_maxnnz, warp_smem_bytes, cF, cD, the_region;
    double time, tstepinv, lambda;
    signed char idxF[100], idxD[100];
};

__host__ __device__ constexpr bool flux_separable(int flux) { return flux == VFVM_NONE || flux == VFVM_FLUX_DIFFUSION || flux == VFVM_FLUX_POWDIFF; }

// species-separable fluxes f_i(u_i,K, u_i,L): 2 partials instead of 2n
template <int FLUX>
__device__ __forceinline__ Dual<2> eval_flux_sep(const double* __restrict__ p, int i, int ns, const Dual<2>& a, const Dual<2>& b) {
    if constexpr (FLUX == VFVM_FLUX_DIFFUSION) return p[i] * (a - b);
    else if constexpr (FLUX == VFVM_FLUX_POWDIFF) return p[i] * (dpowr(a, p[ns]) - dpowr(b, p[ns]));
    else return Dual<2>(0.0);
}

// one off-diagonal block (row r, column L): flux in duals with the reference's orientation, coalesced store of the
// off-diagonal Jacobian block, residual / diagonal-block contributions left in the warp's shared-memory slot kl
template <int NS, int FLUX, bool MULTIREG>
__device__ __forceinline__ void block_entry(const AsmArgs& a, const double* __restrict__ pf, int k, int kl, int r, int L, double fac0, const double* ur,
                                            const double* uc, double* sF, double* sD, bool& nan_seen) {
    const int cF = a.cF;
    const bool pos = r > L;  // row node is edge.node[1] (the larger index): flux(u_row, u_col), sign +
    int64_t it0 = 0, it1 = 1;
    if constexpr (MULTIREG) {
        const int e = a.nz_edge[k];
        it0 = a.ef_colptr[e];
        it1 = a.ef_colptr[e + 1];
    }
    for (int64_t it = it0; it < it1; it++) {
        const bool first = (it == it0);
        double fac = fac0;
        if constexpr (MULTIREG) fac = a.ef_fac[it];
        const double sfac = pos ? fac : -fac;
        if constexpr (flux_separable(FLUX)) {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                Dual<2> x(pos ? ur[i] : uc[i]), y(pos ? uc[i] : ur[i]);
                x.d[0] = 1.0;
                y.d[1] = 1.0;
                const Dual<2> f = eval_flux_sep<FLUX>(pf, i, NS, x, y);
                const double drow = pos ? f.d[0] : f.d[1], dcol = pos ? f.d[1] : f.d[0];
                nan_seen |= (drow != drow) | (dcol != dcol);
                const int p = a.idxF[i * NS + i];
                if (first) {
                    sF[kl * NS + i] = sfac * f.v;
                    if (p >= 0) {
                        sD[kl * cF + p] = sfac * drow;
                        a.offval[(int64_t)p * a.nnz_off + k] = sfac * dcol;
                    }
                } else {
                    sF[kl * NS + i] += sfac * f.v;
                    if (p >= 0) {
                        sD[kl * cF + p] += sfac * drow;
                        a.offval[(int64_t)p * a.nnz_off + k] += sfac * dcol;
                    }
                }
            }
        } else {
            typedef Dual<2 * NS> D;
            D x[NS], y[NS], f[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                x[i] = D(pos ? ur[i] : uc[i]);
                x[i].d[i] = 1.0;
                y[i] = D(pos ? uc[i] : ur[i]);
                y[i].d[NS + i] = 1.0;
                f[i] = D(0.0);
            }
            eval_flux<FLUX, NS>(pf, f, x, y);
#pragma unroll
            for (int i = 0; i < NS; i++) {
                if (first) sF[kl * NS + i] = sfac * f[i].v;
                else sF[kl * NS + i] += sfac * f[i].v;
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    const int p = a.idxF[i * NS + j];
                    if (p < 0) continue;
                    const double drow = pos ? f[i].d[j] : f[i].d[NS + j], dcol = pos ? f[i].d[NS + j] : f[i].d[j];
                    nan_seen |= (drow != drow) | (dcol != dcol);
                    if (first) {
                        sD[kl * cF + p] = sfac * drow;
                        a.offval[(int64_t)p * a.nnz_off + k] = sfac * dcol;
                    } else {
                        sD[kl * cF + p] += sfac * drow;
                        a.offval[(int64_t)p * a.nnz_off + k] += sfac * dcol;
                    }
                }
            }
        }
    }
}

// Warp-autonomous row-group kernel: a warp owns R consecutive node rows per pass (no __syncthreads, no tile table).
//   1. lanes 0..R-1 read their row's [rowptr, rowptr+1) and prefetch the row's node data (u, u_old, node factor, source)
//   2. lane-per-block sweep over the group's contiguous (colidx, nzfac) segment, UNR blocks per lane in flight
//   3. lane-per-row reduction of the shared-memory contributions in column order + node terms, write F / diagonal block
template <int NS, int FLUX, bool MULTIREG, int R, int UNR>
__global__ void __launch_bounds__(ASM_THREADS) k_assemble_rows(const AsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = a.group_maxnnz, cF = a.cF;
    double* sF = (double*)(smem_raw + (size_t)warp * a.warp_smem_bytes);  // M x NS   residual contribution of each block
    double* sD = sF + (size_t)M * NS;                                      // M x cF   diagonal-block contribution of each block
    uint8_t* rowof = (uint8_t*)(sD + (size_t)M * (cF > 0 ? cF : 1));       // M        local row of each block
    const PhysicsDev& ph = *a.ph;
    const double* __restrict__ pf = ph.params + ph.slot[VFVM_SLOT_FLUX].off;
    const bool has_storage = ph.slot[VFVM_SLOT_STORAGE].id != VFVM_NONE;
    bool nan_seen = false;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;

    for (int g = blockIdx.x * wpb + warp; g < a.ngroups; g += nwarps) {
        const int64_t R0 = (int64_t)g * R;
        const int nrows = (int)min((int64_t)R, a.Nown - R0);
        const int myrow = min(lane, nrows - 1);
        const int rb = a.rowptr[R0 + myrow], re = a.rowptr[R0 + myrow + 1];
        const int k0 = __shfl_sync(0xffffffffu, rb, 0), k1 = __shfl_sync(0xffffffffu, re, nrows - 1);
        // prefetch this lane's row data; consumed in the row phase
        const int64_t r = R0 + myrow;
        double u_r[NS], uo_r[NS], src_r[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u_r[i] = a.U[r * NS + i];
            uo_r[i] = has_storage ? a.UOld[r * NS + i] : 0.0;
            src_r[i] = a.src ? a.src[r * NS + i] : 0.0;
        }
        double nfac0 = 0.0;
        if constexpr (!MULTIREG) nfac0 = a.nf_fac[r];
        if (lane < nrows)
            for (int k = rb; k < re; k++) rowof[k - k0] = (uint8_t)lane;
        __syncwarp();

        // ---------------- block phase
        for (int kb = k0; kb < k1; kb += 32 * UNR) {
            int Lc[UNR], rl[UNR];
            double fc[UNR];
            bool act[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = kb + u * 32 + lane;
                act[u] = k < k1;
                Lc[u] = act[u] ? a.colidx[k] : 0;
                fc[u] = (!MULTIREG && act[u]) ? a.nzfac[k] : 0.0;
                rl[u] = act[u] ? rowof[k - k0] : 0;
            }
            double uc[UNR][NS], ur[UNR][NS];
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    uc[u][i] = a.U[(int64_t)Lc[u] * NS + i];
                    ur[u][i] = a.U[(R0 + rl[u]) * NS + i];
                }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = kb + u * 32 + lane;
                if (act[u]) block_entry<NS, FLUX, MULTIREG>(a, pf, k, k - k0, (int)(R0 + rl[u]), Lc[u], fc[u], ur[u], uc[u], sF, sD, nan_seen);
            }
        }
        __syncwarp();

        // ---------------- row phase: lane-per-row segment sums in column order + node terms (K4), write F and diagonal block
        if (lane < nrows) {
            const int kb = rb - k0, ke = re - k0;
            double Fr[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double s = 0.0;
                for (int k = kb; k < ke; k++) s += sF[k * NS + i];
                Fr[i] = s;
            }
            int64_t q0 = r, q1 = r + 1;
            if constexpr (MULTIREG) {
                q0 = a.nf_colptr[r];
                q1 = a.nf_colptr[r + 1];
            }
            typedef Dual<NS> DN;
            DN u[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                u[i] = DN(u_r[i]);
                u[i].d[i] = 1.0;
            }
            for (int64_t q = q0; q < q1; q++) {
                const bool first = (q == q0);
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
                eval_reaction<NS>(ph.slot[VFVM_SLOT_REACTION].id, ph.params + ph.slot[VFVM_SLOT_REACTION].off, rea, u, region);
                if (has_storage) {
                    eval_storage<NS>(ph.slot[VFVM_SLOT_STORAGE].id, ph.params + ph.slot[VFVM_SLOT_STORAGE].off, stor, u);
                    eval_storage<NS>(ph.slot[VFVM_SLOT_STORAGE].id, ph.params + ph.slot[VFVM_SLOT_STORAGE].off, ostor, uo_r);
                }
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    Fr[i] += fac * (rea[i].v - src_r[i] + (stor[i].v - ostor[i]) * a.tstepinv);
#pragma unroll
                    for (int j = 0; j < NS; j++) {
                        const int pD = a.idxD[i * NS + j];
                        if (pD < 0) continue;
                        const double jv = rea[i].d[j] + stor[i].d[j] * a.tstepinv;
                        nan_seen |= (jv != jv);
                        if (first) {
                            double s = 0.0;
                            const int pF = a.idxF[i * NS + j];
                            if (pF >= 0)
                                for (int k = kb; k < ke; k++) s += sD[k * cF + pF];
                            a.diagval[(int64_t)pD * a.Nown + r] = s + jv * fac;
                        } else {
                            a.diagval[(int64_t)pD * a.Nown + r] += jv * fac;
                        }
                    }
                }
            }
            if (q0 == q1) {  // node without any cell: only the flux sums (none) -> zero block
#pragma unroll
                for (int i = 0; i < NS; i++)
#pragma unroll
                    for (int j = 0; j < NS; j++) {
                        const int pD = a.idxD[i * NS + j];
                        if (pD >= 0) a.diagval[(int64_t)pD * a.Nown + r] = 0.0;
                    }
            }
#pragma unroll
            for (int i = 0; i < NS; i++) a.F[r * NS + i] = Fr[i];
        }
        __syncwarp();
    }
    if (nan_seen) atomicOr(a.flags, 1);
}

// ---- fast path: species-separable, antisymmetric flux f_i = D_i (g(u_iK) - g(u_iL)) with diagonal coupling masks ----------
// (linear / power-law diffusion: cfg1, cfg2, cfg3, cfg5).  For these fluxes fac*f and its derivatives are bitwise the same
// for either edge orientation, so no orientation selects are needed; the row lanes seed each block's shared-memory slot
// with the row's unknowns (no row-index table), the block phase overwrites the slot with (residual, diagonal) terms.
template <class T>
__device__ __forceinline__ T reaction_sep(int id, const double* __restrict__ p, int i, int ns, const T& u) {
    switch (id) {
        case VFVM_REACTION_POW: return p[i] * dpowr(u, p[ns + i]);
        case VFVM_REACTION_SINH: return p[i] * (dexp(u) - dexp(-u));
        case VFVM_REACTION_AFFINE: return p[1] + p[0] * u;  // ns == 1 only
        default: return T(0.0);
    }
}
template <class T>
__device__ __forceinline__ T storage_sep(int id, const double* __restrict__ p, int i, int ns, const T& u) {
    switch (id) {
        case VFVM_STORAGE_LINEAR: return p[i] * u;
        case VFVM_STORAGE_POW: return dpowr(p[i] + u, 1.0 / p[ns + i]);
        default: return T(0.0);
    }
}

template <int NS, int FLUX, int R, int UNR>
__global__ void __launch_bounds__(ASM_THREADS) k_assemble_rows_sep(const AsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W = 2 * NS;  // slot: NS residual terms + NS diagonal terms (before: the row's NS unknowns)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* slot = (double*)(smem_raw + (size_t)warp * a.warp_smem_bytes);
    const PhysicsDev& ph = *a.ph;
    const double* __restrict__ pf = ph.params + ph.slot[VFVM_SLOT_FLUX].off;
    const int rid = ph.slot[VFVM_SLOT_REACTION].id, sid = ph.slot[VFVM_SLOT_STORAGE].id;
    const double* __restrict__ pr = ph.params + ph.slot[VFVM_SLOT_REACTION].off;
    const double* __restrict__ ps = ph.params + ph.slot[VFVM_SLOT_STORAGE].off;
    const bool has_storage = sid != VFVM_NONE;
    const int64_t nnz = a.nnz_off;
    bool nan_seen = false;
    const int wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;

    for (int g = blockIdx.x * wpb + warp; g < a.ngroups; g += nwarps) {
        const int64_t R0 = (int64_t)g * R;
        const int nrows = (int)min((int64_t)R, a.Nown - R0);
        const int myrow = min(lane, nrows - 1);
        const int64_t r = R0 + myrow;
        const int rb = a.rowptr[r], re = a.rowptr[r + 1];
        const int k0 = __shfl_sync(0xffffffffu, rb, 0), k1 = __shfl_sync(0xffffffffu, re, nrows - 1);
        double u_r[NS], uo_r[NS], src_r[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u_r[i] = a.U[r * NS + i];
            uo_r[i] = has_storage ? a.UOld[r * NS + i] : 0.0;
            src_r[i] = a.src ? a.src[r * NS + i] : 0.0;
        }
        const double nfac = a.nf_fac[r];
        if (lane < nrows)
            for (int k = rb - k0; k < re - k0; k++) {
#pragma unroll
                for (int i = 0; i < NS; i++) slot[k * W + i] = u_r[i];
            }
        __syncwarp();

        // ---------------- block phase
        for (int kb = k0; kb < k1; kb += 32 * UNR) {
            int Lc[UNR];
            double fc[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = min(kb + u * 32 + lane, k1 - 1);  // clamped: tail lanes recompute the last block (same values)
                Lc[u] = a.colidx[k];
                fc[u] = a.nzfac[k];
            }
            double uc[UNR][NS];
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int i = 0; i < NS; i++) uc[u][i] = a.U[(int64_t)Lc[u] * NS + i];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = kb + u * 32 + lane;
                if (k < k1) {
                    const int kl = k - k0;
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        Dual<2> x(slot[kl * W + i]), y(uc[u][i]);
                        x.d[0] = 1.0;
                        y.d[1] = 1.0;
                        const Dual<2> f = eval_flux_sep<FLUX>(pf, i, NS, x, y);
                        nan_seen |= (f.d[0] != f.d[0]) | (f.d[1] != f.d[1]);
                        slot[kl * W + i] = fc[u] * f.v;
                        slot[kl * W + NS + i] = fc[u] * f.d[0];
                        a.offval[(int64_t)i * nnz + k] = fc[u] * f.d[1];
                    }
                }
            }
        }
        __syncwarp();

        // ---------------- row phase
        if (lane < nrows) {
            double Fr[NS], Dr[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) Fr[i] = Dr[i] = 0.0;
            for (int k = rb - k0; k < re - k0; k++) {
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    Fr[i] += slot[k * W + i];
                    Dr[i] += slot[k * W + NS + i];
                }
            }
#pragma unroll
            for (int i = 0; i < NS; i++) {
                Dual<1> u(u_r[i]);
                u.d[0] = 1.0;
                const Dual<1> rea = reaction_sep(rid, pr, i, NS, u);
                Dual<1> stor(0.0);
                double ostor = 0.0;
                if (has_storage) {
                    stor = storage_sep(sid, ps, i, NS, u);
                    ostor = storage_sep(sid, ps, i, NS, uo_r[i]);
                }
                const double jv = rea.d[0] + stor.d[0] * a.tstepinv;
                nan_seen |= (jv != jv);
                a.F[r * NS + i] = Fr[i] + nfac * (rea.v - src_r[i] + (stor.v - ostor) * a.tstepinv);
                a.diagval[(int64_t)i * a.Nown + r] = Dr[i] + jv * nfac;
            }
        }
        __syncwarp();
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

// picks the block size (in warps) that maximises resident warps per SM for this kernel / per-warp shared memory, then
// launches a persistent grid (one wave: SMs x resident blocks)
struct LaunchPlan {
    int wpb = 0, blocks_per_sm = 0;
    size_t per_warp = 0;
};
template <class Kern>
static void launch_groups(vfvm_handle* h, Kern kern, LaunchPlan& plan, const AsmArgs& a, size_t per_warp) {
    if (plan.wpb == 0 || plan.per_warp != per_warp) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int best_w = 0, best_b = 0;
        for (int w = ASM_WARPS; w >= 1; w--) {
            const size_t smem = per_warp * w;
            if (smem > 227 * 1024) continue;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            int b = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, 32 * w, smem));
            if (b * w > best_b * best_w) {
                best_w = w;
                best_b = b;
            }
        }
        if (best_w == 0) throw std::string("assembly kernel does not fit on an SM: a row group needs ") + std::to_string(per_warp) + " bytes of shared memory per warp";
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * best_w)));
        plan.wpb = best_w;
        plan.blocks_per_sm = best_b;
        plan.per_warp = per_warp;
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::max(1, std::min(cdiv(a.ngroups, plan.wpb), nsm * plan.blocks_per_sm));
    kern<<<grid, 32 * plan.wpb, per_warp * plan.wpb, h->stream>>>(a);
    h->launches++;
}

// ---- host dispatch -------------------------------------------------------------------------------------------
// rows per warp pass / blocks in flight per lane
template <int NS> struct RowCfg { static constexpr int R = 8, UNR = 1; };
template <> struct RowCfg<1> { static constexpr int R = 32, UNR = 4; };
template <> struct RowCfg<2> { static constexpr int R = 16, UNR = 2; };
template <> struct RowCfg<3> { static constexpr int R = 16, UNR = 1; };

int vfvm_rows_per_group(int ns) { return ns == 1 ? 32 : (ns <= 3 ? 16 : 8); }

template <int NS, int FLUX, bool MR>
static void launch_rows_k(vfvm_handle* h, const AsmArgs& a, size_t per_warp) {
    constexpr int R = RowCfg<NS>::R, UNR = RowCfg<NS>::UNR;
    static LaunchPlan plan;
    launch_groups(h, k_assemble_rows<NS, FLUX, MR, R, UNR>, plan, a, per_warp);
}

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

template <int NS, int FLUX>
static void launch_rows_sep(vfvm_handle* h, AsmArgs a) {
    constexpr int R = RowCfg<NS>::R, UNR = RowCfg<NS>::UNR;
    a.warp_smem_bytes = ((h->group_maxnnz * 2 * NS * 8 + 15) / 16) * 16;
    static LaunchPlan plan;
    launch_groups(h, k_assemble_rows_sep<NS, FLUX, R, UNR>, plan, a, (size_t)a.warp_smem_bytes);
}

template <int NS, int FLUX>
static void launch_rows(vfvm_handle* h, const AsmArgs& a, size_t smem) {
    if constexpr (FLUX == VFVM_FLUX_DIFFUSION || FLUX == VFVM_FLUX_POWDIFF) {
        if (fast_path_ok(h)) {
            launch_rows_sep<NS, FLUX>(h, a);
            return;
        }
    }
    if constexpr (flux_supported(FLUX, NS)) {
        if (h->single_region) launch_rows_k<NS, FLUX, false>(h, a, smem);
        else launch_rows_k<NS, FLUX, true>(h, a, smem);
    } else {
        throw std::string("flux id ") + std::to_string(FLUX) + " has no device instantiation for " + std::to_string(NS) + " species";
    }
}

template <int NS>
static void launch_rows_ns(vfvm_handle* h, const AsmArgs& a, size_t smem) {
    switch (h->phys.slot[VFVM_SLOT_FLUX].id) {
        case VFVM_NONE: launch_rows<NS, VFVM_NONE>(h, a, smem); break;
        case VFVM_FLUX_DIFFUSION: launch_rows<NS, VFVM_FLUX_DIFFUSION>(h, a, smem); break;
        case VFVM_FLUX_POWDIFF: launch_rows<NS, VFVM_FLUX_POWDIFF>(h, a, smem); break;
        case VFVM_FLUX_CROSSDIFF2: launch_rows<NS, VFVM_FLUX_CROSSDIFF2>(h, a, smem); break;
        case VFVM_FLUX_SG_UNIPOLAR: launch_rows<NS, VFVM_FLUX_SG_UNIPOLAR>(h, a, smem); break;
        case VFVM_FLUX_SEDAN: launch_rows<NS, VFVM_FLUX_SEDAN>(h, a, smem); break;
        case VFVM_FLUX_SG_BIPOLAR: launch_rows<NS, VFVM_FLUX_SG_BIPOLAR>(h, a, smem); break;
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

// bytes of shared memory one warp needs for a row group
static int warp_smem_bytes(const vfvm_handle* h) {
    const int per = (h->n + std::max(1, h->cF)) * 8 + 1;
    return ((h->group_maxnnz * per + 15) / 16) * 16;
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

int vfvm_assemble_impl(vfvm_handle* h, double time, double tstep, double lambda) {
    cudaStream_t s = h->stream;
    const double tstepinv = 1.0 / tstep;  // src/vfvm_assembly.jl:554 (1/Inf == 0)
    if (tstepinv != 0.0) h->seen_transient = true;
    AsmArgs a;
    a.rowptr = h->rowptr.p;
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
    a.flags = h->flags.p;
    a.nnz_off = h->nnz_off;
    a.Nown = h->Nown;
    a.ngroups = h->ngroups;
    a.group_maxnnz = h->group_maxnnz;
    a.warp_smem_bytes = warp_smem_bytes(h);
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
    const size_t smem = (size_t)a.warp_smem_bytes;  // per warp
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t), s));
    CK(cudaEventRecord(h->ev0, s));
    NS_DISPATCH(h->n, (launch_rows_ns<NS>(h, a, smem)));
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
        b.flags = h->flags.p;
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
    CK(cudaMemcpyAsync(h->flags_host, h->flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev2));
    h->times[VFVM_TIME_ASSEMBLE] = ms;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->times[VFVM_TIME_EDGE_KERNEL] = ms;
    h->precon_valid = false;
    if (h->flags_host[0] & 1) return vfvm_fail(h, VFVM_ERR_NAN, "trying to assemble NaN");
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
