// K4 + K5 + K6: residual + Jacobian assembly (eval_and_assemble, src/vfvm_assembly.jl:520-643).
//
// B200 design (not the reference's edge loop): ONE fused row-tile kernel streams the off-diagonal block CSR
// once.  A CTA owns a tile of consecutive node rows; one thread per off-diagonal block (K,L) evaluates the flux of
// the edge {K,L} in forward-mode duals with the reference's orientation (edge.node[1] = larger node), writes the
// off-diagonal Jacobian block with a fully coalesced store (no edge->nnz scatter, no atomics, no memset of the
// matrix) and leaves its residual / diagonal-block contribution in shared memory; one thread per row then reduces
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
#define ASM_RMAX 256

struct AsmArgs {
    const int32_t* __restrict__ tile_row;
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
    const double* __restrict__ coord;
    double* __restrict__ F;
    double* __restrict__ offval;
    double* __restrict__ diagval;
    const PhysicsDev* __restrict__ ph;
    int32_t* flags;
    int64_t nnz_off, Nown;
    int ntiles, tile_nnz, dim, cF, cD, the_region;
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

template <int NS, int FLUX, bool MULTIREG>
__global__ void __launch_bounds__(ASM_THREADS) k_assemble_rows(const AsmArgs a) {
    extern __shared__ double smem[];
    const int T = a.tile_nnz, cF = a.cF;
    double* sF = smem;                  // T x NS   residual contribution of each block
    double* sD = sF + (size_t)T * NS;   // T x cF   diagonal-block contribution of each block
    int32_t* srp = (int32_t*)(sD + (size_t)T * (cF > 0 ? cF : 1));  // ASM_RMAX+1 row pointers of the tile
    uint8_t* rowof = (uint8_t*)(srp + ASM_RMAX + 1);                 // T: local row of each block
    const int tid = threadIdx.x;
    const double* __restrict__ pf = a.ph->params + a.ph->slot[VFVM_SLOT_FLUX].off;
    bool nan_seen = false;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1], nrows = r1 - r0;
        for (int t = tid; t <= nrows; t += ASM_THREADS) srp[t] = a.rowptr[r0 + t];
        __syncthreads();
        const int k0 = srp[0], k1 = srp[nrows];
        for (int t = tid; t < nrows; t += ASM_THREADS)
            for (int k = srp[t]; k < srp[t + 1]; k++) rowof[k - k0] = (uint8_t)t;
        __syncthreads();

        // ---------------- block phase: one thread per off-diagonal block (row r, column L)
        for (int k = k0 + tid; k < k1; k += ASM_THREADS) {
            const int kl = k - k0;
            const int r = r0 + rowof[kl];
            const int L = a.colidx[k];
            const bool pos = r > L;  // row node is edge.node[1] (the larger index): flux(u_row, u_col), sign +
            int64_t it0 = 0, it1 = 1;
            if constexpr (MULTIREG) {
                const int e = a.nz_edge[k];
                it0 = a.ef_colptr[e];
                it1 = a.ef_colptr[e + 1];
            }
            for (int64_t it = it0; it < it1; it++) {
                const bool first = (it == it0);
                double fac;
                if constexpr (MULTIREG) fac = a.ef_fac[it];
                else fac = a.nzfac[k];
                const double sfac = pos ? fac : -fac;
                if constexpr (flux_separable(FLUX)) {
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        const double ur = a.U[(int64_t)r * NS + i], uc = a.U[(int64_t)L * NS + i];
                        Dual<2> x(pos ? ur : uc), y(pos ? uc : ur);
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
                        const double ur = a.U[(int64_t)r * NS + i], uc = a.U[(int64_t)L * NS + i];
                        x[i] = D(pos ? ur : uc);
                        x[i].d[i] = 1.0;
                        y[i] = D(pos ? uc : ur);
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
        __syncthreads();

        // ---------------- row phase: one thread per node row: segment sums + node terms (K4), write F and diagonal block
        for (int t = tid; t < nrows; t += ASM_THREADS) {
            const int r = r0 + t;
            const int kb = srp[t] - k0, ke = srp[t + 1] - k0;
            const PhysicsDev& ph = *a.ph;
            double Fr[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double s = 0.0;
                for (int k = kb; k < ke; k++) s += sF[k * NS + i];
                Fr[i] = s;
            }
            int64_t q0 = r, q1 = (int64_t)r + 1;
            if constexpr (MULTIREG) {
                q0 = a.nf_colptr[r];
                q1 = a.nf_colptr[r + 1];
            }
            typedef Dual<NS> DN;
            DN u[NS];
            double uo[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                u[i] = DN(a.U[(int64_t)r * NS + i]);
                u[i].d[i] = 1.0;
                uo[i] = a.UOld[(int64_t)r * NS + i];
            }
            for (int64_t q = q0; q < q1; q++) {
                const bool first = (q == q0);
                const double fac = a.nf_fac[q];
                const int region = MULTIREG ? a.nf_region[q] : a.the_region;
                double src[NS], ostor[NS];
                DN rea[NS], stor[NS];
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    src[i] = 0.0;
                    ostor[i] = 0.0;
                    rea[i] = DN(0.0);
                    stor[i] = DN(0.0);
                }
                eval_source<NS>(ph.slot[VFVM_SLOT_SOURCE].id, ph.params + ph.slot[VFVM_SLOT_SOURCE].off, src, a.coord + (int64_t)r * a.dim, a.dim, ph.nodal_source, r);
                eval_reaction<NS>(ph.slot[VFVM_SLOT_REACTION].id, ph.params + ph.slot[VFVM_SLOT_REACTION].off, rea, u, region);
                if (ph.slot[VFVM_SLOT_STORAGE].id != VFVM_NONE) {
                    eval_storage<NS>(ph.slot[VFVM_SLOT_STORAGE].id, ph.params + ph.slot[VFVM_SLOT_STORAGE].off, stor, u);
                    eval_storage<NS>(ph.slot[VFVM_SLOT_STORAGE].id, ph.params + ph.slot[VFVM_SLOT_STORAGE].off, ostor, uo);
                }
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    Fr[i] += fac * (rea[i].v - src[i] + (stor[i].v - ostor[i]) * a.tstepinv);
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
            if (q0 == q1) {  // node without any cell: only the flux sums
#pragma unroll
                for (int i = 0; i < NS; i++)
#pragma unroll
                    for (int j = 0; j < NS; j++) {
                        const int pD = a.idxD[i * NS + j];
                        if (pD >= 0) a.diagval[(int64_t)pD * a.Nown + r] = 0.0;
                    }
            }
#pragma unroll
            for (int i = 0; i < NS; i++) a.F[(int64_t)r * NS + i] = Fr[i];
        }
        __syncthreads();
    }
    if (nan_seen) atomicOr(a.flags, 1);
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

// ---- host dispatch -------------------------------------------------------------------------------------------
template <int NS, int FLUX>
static void launch_rows(vfvm_handle* h, const AsmArgs& a, size_t smem, int grid) {
    if constexpr (flux_supported(FLUX, NS)) {
        if (h->single_region) {
            CK(cudaFuncSetAttribute(k_assemble_rows<NS, FLUX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_assemble_rows<NS, FLUX, false><<<grid, ASM_THREADS, smem, h->stream>>>(a);
        } else {
            CK(cudaFuncSetAttribute(k_assemble_rows<NS, FLUX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_assemble_rows<NS, FLUX, true><<<grid, ASM_THREADS, smem, h->stream>>>(a);
        }
        h->launches++;
    } else {
        throw std::string("flux id ") + std::to_string(FLUX) + " has no device instantiation for " + std::to_string(NS) + " species";
    }
}

template <int NS>
static void launch_rows_ns(vfvm_handle* h, const AsmArgs& a, size_t smem, int grid) {
    switch (h->phys.slot[VFVM_SLOT_FLUX].id) {
        case VFVM_NONE: launch_rows<NS, VFVM_NONE>(h, a, smem, grid); break;
        case VFVM_FLUX_DIFFUSION: launch_rows<NS, VFVM_FLUX_DIFFUSION>(h, a, smem, grid); break;
        case VFVM_FLUX_POWDIFF: launch_rows<NS, VFVM_FLUX_POWDIFF>(h, a, smem, grid); break;
        case VFVM_FLUX_CROSSDIFF2: launch_rows<NS, VFVM_FLUX_CROSSDIFF2>(h, a, smem, grid); break;
        case VFVM_FLUX_SG_UNIPOLAR: launch_rows<NS, VFVM_FLUX_SG_UNIPOLAR>(h, a, smem, grid); break;
        case VFVM_FLUX_SEDAN: launch_rows<NS, VFVM_FLUX_SEDAN>(h, a, smem, grid); break;
        case VFVM_FLUX_SG_BIPOLAR: launch_rows<NS, VFVM_FLUX_SG_BIPOLAR>(h, a, smem, grid); break;
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

size_t vfvm_asm_smem(const vfvm_handle* h) {
    return (size_t)h->tile_nnz * (h->n + std::max(1, h->cF)) * sizeof(double) + (ASM_RMAX + 1) * sizeof(int32_t) + (size_t)h->tile_nnz;
}

int vfvm_assemble_impl(vfvm_handle* h, double time, double tstep, double lambda) {
    cudaStream_t s = h->stream;
    const double tstepinv = 1.0 / tstep;  // src/vfvm_assembly.jl:554 (1/Inf == 0)
    if (tstepinv != 0.0) h->seen_transient = true;
    AsmArgs a;
    a.tile_row = h->tile_row.p;
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
    a.coord = h->coord.p;
    a.F = h->vec[VFVM_VEC_RESIDUAL].p;
    a.offval = h->offval.p;
    a.diagval = h->diagval.p;
    a.ph = h->phys_dev.p;
    a.flags = h->flags.p;
    a.nnz_off = h->nnz_off;
    a.Nown = h->Nown;
    a.ntiles = h->ntiles;
    a.tile_nnz = h->tile_nnz;
    a.dim = h->dim;
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
    const size_t smem = vfvm_asm_smem(h);
    const int grid = h->ntiles;
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t), s));
    CK(cudaEventRecord(h->ev0, s));
    NS_DISPATCH(h->n, (launch_rows_ns<NS>(h, a, smem, grid)));
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
