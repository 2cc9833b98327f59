// Post-processing integrals on the device: the node and edge loops of the assembly without AD.
//   integrate(system, F, U)      src/vfvm_postprocess.jl:18-67    -> vfvm_integrate      (F: registered reaction / storage function or the identity)
//   edgeintegrate(system, F, U)  src/vfvm_postprocess.jl:109-146  -> vfvm_edgeintegrate  (F: registered flux or the W^{1,p} integrand of :300-312)
// Both return the n x ncellregions matrix of region-wise integrals of the resident vector `which`.  One pass per cell region;
// block partial sums in a fixed order + a single-block finalize => bitwise reproducible.  With several ranks every rank
// integrates its owned nodes, an edge cut by a partition boundary counts half on either side, and the sums are all-reduced.
#include "physics.cuh"
#include "vfvm_internal.h"

int vfvm_comm_allreduce_sum(vfvm_handle* h, double* dev, int count);

namespace {

struct FnArgs {
    int slot, id, np;
    double p[VFVM_MAX_PARAMS];
};

__device__ __forceinline__ double pp_block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += sh[i];
    return r;
}

template <int NS>
__global__ void k_integrate_nodes(int64_t Nown, int region, const int64_t* __restrict__ colptr, const int32_t* __restrict__ nregion, const double* __restrict__ nfac,
                                  const double* __restrict__ U, const FnArgs fn, unsigned rsm, double* __restrict__ part) {
    // rsm: species enabled in this cell region (assemble_res goes through isregionspecies, src/vfvm_assemblydata.jl:310-332)
    __shared__ double red[32];
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) acc[i] = 0.0;
    for (int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; K < Nown; K += (int64_t)gridDim.x * blockDim.x) {
        double fac = 0.0;
        for (int64_t q = colptr[K]; q < colptr[K + 1]; q++)
            if (nregion[q] == region) fac += nfac[q];  // one item per (node, region)
        if (fac == 0.0) continue;
        double u[NS], f[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u[i] = U[K * NS + i];
            f[i] = 0.0;
        }
        if (fn.id == VFVM_NONE) {
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = u[i];
        } else if (fn.slot == VFVM_SLOT_STORAGE) {
            eval_storage<NS>(fn.id, fn.p, f, u);
        } else {
            eval_reaction<NS>(fn.id, fn.p, f, u, region);
        }
#pragma unroll
        for (int i = 0; i < NS; i++)
            if ((rsm >> i) & 1u) acc[i] += fac * f[i];
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double s = pp_block_sum(acc[i], red);
        if (threadIdx.x == 0) part[(int64_t)i * gridDim.x + blockIdx.x] = s;
    }
}

// integrate(system, F, U; boundary = true), src/vfvm_postprocess.jl:29-46: one thread per boundary node over its (bface, local node)
// items of boundary region `bregion`; the registered node / boundary functions depend on the item only through its region
template <int NS>
__global__ void k_integrate_bnodes(int64_t nbnodes, int bregion, int dim, const int32_t* __restrict__ bn_node, const int32_t* __restrict__ bn_ptr,
                                   const int32_t* __restrict__ bn_bface, const int32_t* __restrict__ bn_local, const int32_t* __restrict__ bfaceregions,
                                   const double* __restrict__ bfnf, const double* __restrict__ U, const FnArgs fn, const int32_t* __restrict__ node_active,
                                   double* __restrict__ part) {
    __shared__ double red[32];
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) acc[i] = 0.0;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbnodes; b += (int64_t)gridDim.x * blockDim.x) {
        double fac = 0.0;
        for (int q = bn_ptr[b]; q < bn_ptr[b + 1]; q++) {
            const int ibf = bn_bface[q];
            if (bfaceregions[ibf] == bregion) fac += bfnf[(int64_t)ibf * dim + bn_local[q]];
        }
        if (fac == 0.0) continue;
        const int K = bn_node[b];
        double u[NS], f[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            u[i] = U[(int64_t)K * NS + i];
            f[i] = 0.0;
        }
        if (fn.id == VFVM_NONE) {
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = u[i];
        } else if (fn.slot == VFVM_SLOT_BREACTION) {
            eval_breaction_fn<NS>(fn.id, fn.p, f, u, bregion);
        } else if (fn.slot == VFVM_SLOT_BSTORAGE) {
            eval_bstorage_fn<NS>(fn.id, fn.p, f, u, bregion);
        } else if (fn.slot == VFVM_SLOT_STORAGE) {
            eval_storage<NS>(fn.id, fn.p, f, u);
        } else {
            eval_reaction<NS>(fn.id, fn.p, f, u, bregion);
        }
        const unsigned act = node_active ? (unsigned)node_active[K] : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < NS; i++)
            if ((act >> i) & 1u) acc[i] += fac * f[i];
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double s = pp_block_sum(acc[i], red);
        if (threadIdx.x == 0) part[(int64_t)i * gridDim.x + blockIdx.x] = s;
    }
}

// FLUX == -1: dim ((u_K - u_L) / h)^p;  FLUX == -2: (u_K + u_L) / 2
template <int NS, int FLUX>
__global__ void k_integrate_edges(int64_t E, int64_t Nown, int region, int dim, const int32_t* __restrict__ edgenodes, const int64_t* __restrict__ colptr,
                                  const int32_t* __restrict__ eregion, const double* __restrict__ efac, const double* __restrict__ coord, const double* __restrict__ U,
                                  const FnArgs fn, unsigned rsm, const int32_t* __restrict__ node_active, double* __restrict__ part) {
    // a species takes part on an edge of this region if it is enabled in the region and defined at both end nodes
    // (assemble_res(::Edge), src/vfvm_assemblydata.jl:385-405); node_active == null: every species everywhere
    __shared__ double red[32];
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) acc[i] = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        double fac = 0.0;
        for (int64_t q = colptr[e]; q < colptr[e + 1]; q++)
            if (eregion[q] == region) fac += efac[q];
        const int K = edgenodes[2 * e], L = edgenodes[2 * e + 1];
        const double wgt = 0.5 * ((K < Nown ? 1.0 : 0.0) + (L < Nown ? 1.0 : 0.0));  // partition boundaries: half on either side
        if (fac == 0.0 || wgt == 0.0) continue;
        double h2 = 0.0;
        for (int d = 0; d < dim; d++) {
            const double dx = coord[(int64_t)K * dim + d] - coord[(int64_t)L * dim + d];
            h2 += dx * dx;
        }
        const double hh = sqrt(h2);
        double uK[NS], uL[NS], f[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            uK[i] = U[(int64_t)K * NS + i];
            uL[i] = U[(int64_t)L * NS + i];
            f[i] = 0.0;
        }
        if constexpr (FLUX == -1) {
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = dim * dpowr((uK[i] - uL[i]) / hh, fn.p[0]);
        } else if constexpr (FLUX == -2) {  // edge average (test/test120_norms.jl:35-38)
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = 0.5 * (uK[i] + uL[i]);
        } else {
            eval_flux<FLUX, NS>(fn.p, f, uK, uL);
        }
        const unsigned act = node_active ? (rsm & (unsigned)node_active[K] & (unsigned)node_active[L]) : rsm;
#pragma unroll
        for (int i = 0; i < NS; i++)
            if ((act >> i) & 1u) acc[i] += wgt * (hh * hh * fac * f[i] / dim);
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double s = pp_block_sum(acc[i], red);
        if (threadIdx.x == 0) part[(int64_t)i * gridDim.x + blockIdx.x] = s;
    }
}

// mass_matrix(state), src/vfvm_diffeq_interface.jl:60-101: storage Jacobian at U = 0 times the node factors, one n x n block per node
template <int NS>
__global__ void k_mass_matrix(int64_t Nown, const int64_t* __restrict__ colptr, const int32_t* __restrict__ nregion, const double* __restrict__ nfac,
                              const PhysicsDev* __restrict__ ph, const unsigned short* __restrict__ rsmask, double* __restrict__ out) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= Nown) return;
    typedef Dual<NS> D;
    D u[NS], st[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) {
        u[i] = D(0.0);
        u[i].d[i] = 1.0;
        st[i] = D(0.0);
    }
    eval_storage<NS>(ph->slot[VFVM_SLOT_STORAGE].id, ph->params + ph->slot[VFVM_SLOT_STORAGE].off, st, u);  // the registered storages do not depend on the region
    double M[NS * NS];
#pragma unroll
    for (int i = 0; i < NS * NS; i++) M[i] = 0.0;
    for (int64_t q = colptr[K]; q < colptr[K + 1]; q++) {
        const unsigned m = rsmask[nregion[q] - 1];
        const double fac = nfac[q];
#pragma unroll
        for (int i = 0; i < NS; i++)
#pragma unroll
            for (int j = 0; j < NS; j++)
                if (((m >> i) & 1u) && ((m >> j) & 1u)) M[i * NS + j] += st[i].d[j] * fac;
    }
#pragma unroll
    for (int i = 0; i < NS * NS; i++) out[K * NS * NS + i] = M[i];
}

// flux callback of every edge without form factor (edge loop of nodeflux, src/vfvm_postprocess.jl:191-207): out[e*NS + i]
template <int NS, int FLUX>
__global__ void k_edge_flux(int64_t E, const int32_t* __restrict__ edgenodes, const double* __restrict__ U, const FnArgs fn, double* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int K = edgenodes[2 * e], L = edgenodes[2 * e + 1];
    double uK[NS], uL[NS], f[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) {
        uK[i] = U[(int64_t)K * NS + i];
        uL[i] = U[(int64_t)L * NS + i];
        f[i] = 0.0;
    }
    eval_flux<FLUX, NS>(fn.p, f, uK, uL);
#pragma unroll
    for (int i = 0; i < NS; i++) out[e * NS + i] = f[i];
}

__global__ void k_pp_finalize(const double* __restrict__ part, int nparts, int nvals, double* __restrict__ out) {
    __shared__ double red[32];
    for (int v = 0; v < nvals; v++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[(int64_t)v * nparts + i];
        const double r = pp_block_sum(s, red);
        if (threadIdx.x == 0) out[v] = r;
        __syncthreads();
    }
}

const int PP_GRID = 148 * 4, PP_THREADS = 256;

#define PP_NS(n, ...)                                                                                                     \
    switch (n) {                                                                                                          \
        case 1: { constexpr int NS = 1; __VA_ARGS__; } break;                                                             \
        case 2: { constexpr int NS = 2; __VA_ARGS__; } break;                                                             \
        case 3: { constexpr int NS = 3; __VA_ARGS__; } break;                                                             \
        case 4: { constexpr int NS = 4; __VA_ARGS__; } break;                                                             \
        case 5: { constexpr int NS = 5; __VA_ARGS__; } break;                                                             \
        case 10: { constexpr int NS = 10; __VA_ARGS__; } break;                                                           \
        default: throw std::string("number of species without device instantiation (supported: 1,2,3,4,5,10)");         \
    }

template <int NS, int FLUX>
void launch_edges(vfvm_handle* h, int region, const double* U, const FnArgs& fn, unsigned rsm, double* part) {
    if constexpr (FLUX < 0 || flux_supported(FLUX, NS)) {
        k_integrate_edges<NS, FLUX><<<PP_GRID, PP_THREADS, 0, h->stream>>>(h->E, h->Nown, region, h->dim, h->edgenodes.p, h->ef_colptr.p, h->ef_region.p, h->ef_fac.p,
                                                                           h->coord.p, U, fn, rsm, h->masked ? h->node_active.p : nullptr, part);
    } else {
        throw std::string("flux id ") + std::to_string(FLUX) + " has no device instantiation for " + std::to_string(NS) + " species";
    }
}
template <int NS>
void launch_edges_ns(vfvm_handle* h, int region, const double* U, const FnArgs& fn, unsigned rsm, double* part) {
    switch (fn.id) {
        case -1: launch_edges<NS, -1>(h, region, U, fn, rsm, part); break;
        case -2: launch_edges<NS, -2>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_DIFFUSION: launch_edges<NS, VFVM_FLUX_DIFFUSION>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_POWDIFF: launch_edges<NS, VFVM_FLUX_POWDIFF>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_CROSSDIFF2: launch_edges<NS, VFVM_FLUX_CROSSDIFF2>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_SG_UNIPOLAR: launch_edges<NS, VFVM_FLUX_SG_UNIPOLAR>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_SEDAN: launch_edges<NS, VFVM_FLUX_SEDAN>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_SG_BIPOLAR: launch_edges<NS, VFVM_FLUX_SG_BIPOLAR>(h, region, U, fn, rsm, part); break;
        case VFVM_FLUX_MIXTURE: launch_edges<NS, VFVM_FLUX_MIXTURE>(h, region, U, fn, rsm, part); break;
        default: throw std::string("unregistered flux id");
    }
}

template <int NS, int FLUX>
void launch_edge_flux(vfvm_handle* h, const double* U, const FnArgs& fn, double* out) {
    if constexpr (flux_supported(FLUX, NS)) {
        k_edge_flux<NS, FLUX><<<cdiv(h->E, 256), 256, 0, h->stream>>>(h->E, h->edgenodes.p, U, fn, out);
    } else {
        throw std::string("flux id ") + std::to_string(FLUX) + " has no device instantiation for " + std::to_string(NS) + " species";
    }
}
template <int NS>
void launch_edge_flux_ns(vfvm_handle* h, const double* U, const FnArgs& fn, double* out) {
    switch (fn.id) {
        case VFVM_FLUX_DIFFUSION: launch_edge_flux<NS, VFVM_FLUX_DIFFUSION>(h, U, fn, out); break;
        case VFVM_FLUX_POWDIFF: launch_edge_flux<NS, VFVM_FLUX_POWDIFF>(h, U, fn, out); break;
        case VFVM_FLUX_CROSSDIFF2: launch_edge_flux<NS, VFVM_FLUX_CROSSDIFF2>(h, U, fn, out); break;
        case VFVM_FLUX_SG_UNIPOLAR: launch_edge_flux<NS, VFVM_FLUX_SG_UNIPOLAR>(h, U, fn, out); break;
        case VFVM_FLUX_SEDAN: launch_edge_flux<NS, VFVM_FLUX_SEDAN>(h, U, fn, out); break;
        case VFVM_FLUX_SG_BIPOLAR: launch_edge_flux<NS, VFVM_FLUX_SG_BIPOLAR>(h, U, fn, out); break;
        case VFVM_FLUX_MIXTURE: launch_edge_flux<NS, VFVM_FLUX_MIXTURE>(h, U, fn, out); break;
        default: throw std::string("unregistered flux id");
    }
}

int integrate_impl(vfvm_handle* h, bool edges, int slot, int id, const double* params, int np, int which, double* out) {
    if (np < 0 || np > VFVM_MAX_PARAMS) return vfvm_fail(h, VFVM_ERR_ARG, "too many parameters");
    FnArgs fn;
    memset(&fn, 0, sizeof(fn));
    fn.slot = slot;
    fn.id = id;
    fn.np = np;
    for (int i = 0; i < np; i++) fn.p[i] = params[i];
    const int n = h->n, nreg = h->ncellregions;
    DevBuf<double> part, res;
    part.alloc((size_t)n * PP_GRID);
    res.alloc((size_t)n * nreg);
    const double* U = h->vec[which].p;
    for (int r = 1; r <= nreg; r++) {
        unsigned rsm = 0;
        for (int i = 0; i < n; i++)
            if (h->region_species[(size_t)(r - 1) * n + i]) rsm |= 1u << i;
        if (edges) {
            PP_NS(n, (launch_edges_ns<NS>(h, r, U, fn, rsm, part.p)));
        } else {
            PP_NS(n, (k_integrate_nodes<NS><<<PP_GRID, PP_THREADS, 0, h->stream>>>(h->Nown, r, h->nf_colptr.p, h->nf_region.p, h->nf_fac.p, U, fn, rsm, part.p)));
        }
        k_pp_finalize<<<1, 1024, 0, h->stream>>>(part.p, PP_GRID, n, res.p + (size_t)(r - 1) * n);
        h->launches += 2;
    }
    for (int off = 0; off < n * nreg; off += VFVM_PEER_RED_W) vfvm_comm_allreduce_sum(h, res.p + off, std::min(VFVM_PEER_RED_W, n * nreg - off));
    CK(cudaMemcpyAsync(out, res.p, sizeof(double) * n * nreg, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return vfvm_peer_check(h);
}

}  // namespace

extern "C" int vfvm_integrate(vfvm_handle* h, int slot, int id, const double* params, int np, int which, double* out) {
    if (!h || !h->have_geometry || !h->have_system || !out || which < 0 || which > 3) return vfvm_fail(h, VFVM_ERR_ARG, "vfvm_integrate: geometry and system first; vector id 0..3");
    if (slot != VFVM_SLOT_REACTION && slot != VFVM_SLOT_STORAGE) return vfvm_fail(h, VFVM_ERR_ARG, "node functions are registered reaction or storage functions");
    if (!h->vec[which].p) return vfvm_fail(h, VFVM_ERR_STATE, "vector not set (vfvm_build_pattern allocates the resident vectors)");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        return integrate_impl(h, false, slot, id, params, np, which, out);
    })
}

extern "C" int vfvm_integrate_boundary(vfvm_handle* h, int slot, int id, const double* params, int np, int which, double* out) {
    if (!h || !h->have_pattern || !out || which < 0 || which > 3 || np < 0 || np > VFVM_MAX_PARAMS)
        return vfvm_fail(h, VFVM_ERR_ARG, "vfvm_integrate_boundary: pattern first (the boundary-node lists are built there); vector id 0..3");
    if (!(slot == VFVM_SLOT_BREACTION || slot == VFVM_SLOT_REACTION || slot == VFVM_SLOT_STORAGE || slot == VFVM_SLOT_BSTORAGE))
        return vfvm_fail(h, VFVM_ERR_ARG, "boundary integrals take a registered boundary reaction / boundary storage / reaction / storage function");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        FnArgs fn;
        memset(&fn, 0, sizeof(fn));
        fn.slot = slot;
        fn.id = id;
        fn.np = np;
        for (int i = 0; i < np; i++) fn.p[i] = params[i];
        const int n = h->n, nreg = std::max(1, h->phys.nbregions);
        DevBuf<double> part, res;
        part.alloc((size_t)n * PP_GRID);
        res.alloc((size_t)n * nreg);
        CK(cudaMemsetAsync(res.p, 0, sizeof(double) * n * nreg, h->stream));
        for (int r = 1; r <= nreg && h->nbnodes; r++) {
            PP_NS(n, (k_integrate_bnodes<NS><<<PP_GRID, PP_THREADS, 0, h->stream>>>(h->nbnodes, r, h->dim, h->bn_node.p, h->bn_ptr.p, h->bn_bface.p, h->bn_local.p, h->bfaceregions.p,
                                                                                    h->bfacenodefac.p, h->vec[which].p, fn, h->masked ? h->node_active.p : nullptr, part.p)));
            k_pp_finalize<<<1, 1024, 0, h->stream>>>(part.p, PP_GRID, n, res.p + (size_t)(r - 1) * n);
            h->launches += 2;
        }
        for (int off = 0; off < n * nreg; off += VFVM_PEER_RED_W) vfvm_comm_allreduce_sum(h, res.p + off, std::min(VFVM_PEER_RED_W, n * nreg - off));
        CK(cudaMemcpyAsync(out, res.p, sizeof(double) * n * nreg, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        return vfvm_peer_check(h);
    })
}

extern "C" int vfvm_edgeflux(vfvm_handle* h, int id, const double* params, int np, int which, double* out) {
    if (!h || !h->have_geometry || !h->have_system || !out || which < 0 || which > 3 || np < 0 || np > VFVM_MAX_PARAMS)
        return vfvm_fail(h, VFVM_ERR_ARG, "vfvm_edgeflux: geometry and system first; vector id 0..3");
    if (!h->vec[which].p) return vfvm_fail(h, VFVM_ERR_STATE, "vector not set (vfvm_build_pattern allocates the resident vectors)");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        FnArgs fn;
        memset(&fn, 0, sizeof(fn));
        fn.slot = VFVM_SLOT_FLUX;
        fn.id = id;
        fn.np = np;
        for (int i = 0; i < np; i++) fn.p[i] = params[i];
        DevBuf<double> res;
        res.alloc((size_t)h->n * std::max<int64_t>(1, h->E));
        if (h->E) PP_NS(h->n, (launch_edge_flux_ns<NS>(h, h->vec[which].p, fn, res.p)));
        h->launches++;
        CK(cudaMemcpyAsync(out, res.p, sizeof(double) * h->n * h->E, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
    })
    return VFVM_OK;
}

extern "C" int vfvm_mass_matrix(vfvm_handle* h, double* out) {
    if (!h || !h->have_geometry || !h->have_system || !out) return vfvm_fail(h, VFVM_ERR_ARG, "vfvm_mass_matrix: geometry and system first");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        vfvm_sync_physics(h);
        const int n = h->n;
        std::vector<unsigned short> rs(VFVM_MAX_CREGIONS, 0);
        for (int r = 0; r < h->ncellregions; r++)
            for (int i = 0; i < n; i++)
                if (h->region_species[(size_t)r * n + i]) rs[r] |= (unsigned short)(1u << i);
        DevBuf<unsigned short> drs;
        drs.upload(rs.data(), rs.size(), h->stream);
        DevBuf<double> res;
        res.alloc((size_t)n * n * h->Nown);
        PP_NS(n, (k_mass_matrix<NS><<<cdiv(h->Nown, 128), 128, 0, h->stream>>>(h->Nown, h->nf_colptr.p, h->nf_region.p, h->nf_fac.p, h->phys_dev.p, drs.p, res.p)));
        h->launches++;
        CK(cudaMemcpyAsync(out, res.p, sizeof(double) * n * n * h->Nown, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
    })
    return VFVM_OK;
}

extern "C" int vfvm_edgeintegrate(vfvm_handle* h, int id, const double* params, int np, int which, double* out) {
    if (!h || !h->have_geometry || !h->have_system || !out || which < 0 || which > 3) return vfvm_fail(h, VFVM_ERR_ARG, "vfvm_edgeintegrate: geometry and system first; vector id 0..3");
    if (!h->vec[which].p) return vfvm_fail(h, VFVM_ERR_STATE, "vector not set (vfvm_build_pattern allocates the resident vectors)");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        return integrate_impl(h, true, VFVM_SLOT_FLUX, id, params, np, which, out);
    })
}
