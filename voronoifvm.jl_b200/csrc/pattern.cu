// K3: sparsity pattern + maps.  Replaces what rawupdateindex!/flush! of ExtendableSparse build implicitly during
// the first assembly (src/vfvm_assembly.jl:26, src/vfvm_solver.jl:242).
//
// Device layout ("DBSR": diagonal + off-diagonal block planes, off-diagonal part in sliced ELLPACK order, slice = 32 rows):
//   rowptr[Nown+1]                    CSR row offsets of the off-diagonal node graph (row lengths; export / ILU0)
//   sell_ptr[nslices+1]               first entry of each slice; entry j of row r lives at sell_ptr[r/32] + 32*j + r%32, so that
//                                     a warp working lane-per-row reads colidx / nzfac and writes the Jacobian fully coalesced
//   colidx[nnz_sell]                  column (neighbour node) per entry, ascending within a row; padding entries point to the
//                                     row itself and carry a zero form factor (they assemble / multiply to exact zeros)
//   nzfac[nnz_sell]                   edge form factor per entry (summed over the cell regions of the edge); nz_edge: edge id (-1 = padding)
//   offval[cF][nnz_sell]              one plane per (i,j) of the flux species-coupling mask
//   diagval[cD][Nown]                 one plane per (i,j) of the diagonal-block mask
// The scalar CSR/CSC pattern the reference would hold (value-dependent through _addnz, src/vfvm_assembly.jl:21-28)
// is derived from the block pattern and the per-physics coupling masks by the getters below.
#include <algorithm>
#include <cub/cub.cuh>

#include "vfvm_internal.h"

namespace {

__global__ void k_dir_keys(int64_t E, int64_t Nown, const int32_t* __restrict__ edgenodes, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint64_t hi = (uint32_t)edgenodes[2 * e], lo = (uint32_t)edgenodes[2 * e + 1];
    keys[2 * e] = ((int64_t)hi < Nown) ? ((hi << 32) | lo) : ~0ull;      // rows of halo nodes are not stored: sorted to the end
    keys[2 * e + 1] = ((int64_t)lo < Nown) ? ((lo << 32) | hi) : ~0ull;
    vals[2 * e] = (int32_t)e;
    vals[2 * e + 1] = (int32_t)e;
}

__global__ void k_rowptr(int64_t nrows, int64_t nkeys, const uint64_t* __restrict__ keys, int32_t* __restrict__ rowptr) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K > nrows) return;
    int64_t lo = 0, hi = nkeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> 32) < K) lo = mid + 1;
        else hi = mid;
    }
    rowptr[K] = (int32_t)lo;
}

// slice width = longest row of the slice
__global__ void k_slice_width(int nslices, int64_t Nown, const int32_t* __restrict__ rowptr, int64_t* __restrict__ sell_ptr) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= nslices) return;
    const int64_t r = (int64_t)g * 32 + lane;
    int len = r < Nown ? rowptr[r + 1] - rowptr[r] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) {
        sell_ptr[g + 1] = (int64_t)len * 32;  // exclusive scan follows
        if (g == 0) sell_ptr[0] = 0;
    }
}

// CSR (sorted keys) -> SELL-32 entries, one warp per slice, lane per row
__global__ void k_fill_sell(int nslices, int64_t Nown, const int32_t* __restrict__ rowptr, const int64_t* __restrict__ sell_ptr64,
                            const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals, const int64_t* __restrict__ ef_colptr,
                            const double* __restrict__ ef_fac,
                            int32_t* __restrict__ sell_ptr, int32_t* __restrict__ colidx, int32_t* __restrict__ nz_edge, double* __restrict__ nzfac,
                            int32_t* __restrict__ lowlen) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= nslices) return;
    const int64_t r = (int64_t)g * 32 + lane;
    const int64_t base = sell_ptr64[g];
    const int w = (int)((sell_ptr64[g + 1] - base) >> 5);
    if (lane == 0) {
        sell_ptr[g] = (int32_t)base;
        if (g == nslices - 1) sell_ptr[g + 1] = (int32_t)sell_ptr64[g + 1];
    }
    const int rb = r < Nown ? rowptr[r] : 0, len = r < Nown ? rowptr[r + 1] - rb : 0;
    const int32_t self = (int32_t)min(r, Nown - 1);
    int nlow = 0;
    for (int j = 0; j < w; j++) {
        const int64_t e = base + (int64_t)j * 32 + lane;
        if (j < len) {
            const int32_t c = (int32_t)(keys[rb + j] & 0xffffffffu);
            const int32_t ed = vals[rb + j];
            colidx[e] = c;
            nz_edge[e] = ed;
            // an edge on a region interface carries one factor per adjacent cell region; no registered flux depends on the
            // region, so the factors are summed here (in region order) and the flux is evaluated once per edge end
            double fsum = 0.0;
            for (int64_t q = ef_colptr[ed]; q < ef_colptr[ed + 1]; q++) fsum += ef_fac[q];
            nzfac[e] = fsum;
            nlow += (c < r) ? 1 : 0;
        } else {
            colidx[e] = self;
            nz_edge[e] = -1;
            nzfac[e] = 0.0;
        }
    }
    if (r < Nown) lowlen[r] = nlow;
}

}  // namespace

// species-coupling masks of the registered physics; a coefficient that is exactly zero removes the coupling because
// the reference never inserts Jacobian entries whose value is exactly zero
int vfvm_physics_masks(vfvm_handle* h) {
    const int n = h->n;
    Masks& m = h->masks;
    memset(&m, 0, sizeof(m));
    const PhysicsDev& ph = h->phys;
    auto P = [&](int slot) { return ph.params + ph.slot[slot].off; };
    auto full = [&](uint64_t* mk) {
        for (int i = 0; i < n * n; i++) mask_set(mk, i);
    };
    {
        const double* p = P(VFVM_SLOT_FLUX);
        switch (ph.slot[VFVM_SLOT_FLUX].id) {
            case VFVM_NONE: break;
            case VFVM_FLUX_DIFFUSION:
            case VFVM_FLUX_POWDIFF:
                for (int i = 0; i < n; i++)
                    if (p[i] != 0.0) mask_set(m.flux, i * n + i);
                break;
            case VFVM_FLUX_CROSSDIFF2:
            case VFVM_FLUX_MIXTURE: full(m.flux); break;
            case VFVM_FLUX_SG_UNIPOLAR:
            case VFVM_FLUX_SEDAN: {
                const bool sedan = ph.slot[VFVM_SLOT_FLUX].id == VFVM_FLUX_SEDAN;
                const int iphi = (int)p[sedan ? 2 : 1], ic = (int)p[sedan ? 3 : 2];
                mask_set(m.flux, iphi * n + iphi);
                mask_set(m.flux, ic * n + iphi);
                mask_set(m.flux, ic * n + ic);
                break;
            }
            case VFVM_FLUX_SG_BIPOLAR:
                mask_set(m.flux, 2 * n + 2);
                mask_set(m.flux, 0 * n + 0);
                mask_set(m.flux, 0 * n + 2);
                mask_set(m.flux, 1 * n + 1);
                mask_set(m.flux, 1 * n + 2);
                break;
            default: return VFVM_ERR_UNREGISTERED;
        }
    }
    {
        const double* p = P(VFVM_SLOT_REACTION);
        switch (ph.slot[VFVM_SLOT_REACTION].id) {
            case VFVM_NONE: break;
            case VFVM_REACTION_POW:
            case VFVM_REACTION_SINH:
                for (int i = 0; i < n; i++)
                    if (p[i] != 0.0) mask_set(m.reaction, i * n + i);
                break;
            case VFVM_REACTION_AFFINE:
                for (int i = 0; i < n * n; i++)
                    if (p[i] != 0.0) mask_set(m.reaction, i);
                break;
            case VFVM_REACTION_REGION_AFFINE: {
                const int nreg = (int)p[0];
                for (int r = 0; r < nreg; r++)
                    for (int i = 0; i < n * n; i++)
                        if (p[1 + r * (n * n + n) + i] != 0.0) mask_set(m.reaction, i);
                break;
            }
            case VFVM_REACTION_BILINEAR2:
            case VFVM_REACTION_BIPOLAR: full(m.reaction); break;
            default: return VFVM_ERR_UNREGISTERED;
        }
    }
    {
        const double* p = P(VFVM_SLOT_STORAGE);
        switch (ph.slot[VFVM_SLOT_STORAGE].id) {
            case VFVM_NONE: break;
            case VFVM_STORAGE_LINEAR:
                for (int i = 0; i < n; i++)
                    if (p[i] != 0.0) mask_set(m.storage, i * n + i);
                break;
            case VFVM_STORAGE_POW:
                for (int i = 0; i < n; i++) mask_set(m.storage, i * n + i);
                break;
            case VFVM_STORAGE_BIPOLAR:
                mask_set(m.storage, 0 * n + 0);
                mask_set(m.storage, 0 * n + 2);
                mask_set(m.storage, 1 * n + 1);
                mask_set(m.storage, 1 * n + 2);
                break;
            default: return VFVM_ERR_UNREGISTERED;
        }
    }
    switch (ph.slot[VFVM_SLOT_SOURCE].id) {
        case VFVM_NONE:
        case VFVM_SOURCE_CONST:
        case VFVM_SOURCE_GAUSS:
        case VFVM_SOURCE_XSINYEXPZ:
        case VFVM_SOURCE_STEP1D:
        case VFVM_SOURCE_AFFINE_X:
        case VFVM_SOURCE_NODAL: break;
        default: return VFVM_ERR_UNREGISTERED;
    }
    switch (ph.slot[VFVM_SLOT_BREACTION].id) {
        case VFVM_NONE: break;
        case VFVM_BREACTION_LINEAR: {
            const double* p = P(VFVM_SLOT_BREACTION);
            for (int i = 0; i < n * n; i++)
                if (p[1 + i] != 0.0) mask_set(m.boundary, i);
            break;
        }
        case VFVM_BREACTION_POW: {
            const double* p = P(VFVM_SLOT_BREACTION);
            for (int i = 0; i < n; i++)
                if (p[1 + i] != 0.0) mask_set(m.boundary, i * n + i);
            break;
        }
        case VFVM_BREACTION_CATALYSIS: {  // f_A(u_A, u_C), f_B(u_B, u_C), f_C(u_A, u_B, u_C)
            const double* p = P(VFVM_SLOT_BREACTION);
            const int iA = (int)p[6], iB = (int)p[7], iC = (int)p[8];
            const int pairs[7][2] = {{iA, iA}, {iA, iC}, {iB, iB}, {iB, iC}, {iC, iA}, {iC, iB}, {iC, iC}};
            for (auto& q : pairs) mask_set(m.boundary, q[0] * n + q[1]);
            break;
        }
        default: return VFVM_ERR_UNREGISTERED;
    }
    switch (ph.slot[VFVM_SLOT_BSTORAGE].id) {
        case VFVM_NONE: break;
        case VFVM_BSTORAGE_LINEAR: {
            const double* p = P(VFVM_SLOT_BSTORAGE);
            for (int i = 0; i < n; i++)
                if (p[1 + i] != 0.0) mask_set(m.boundary, i * n + i);
            break;
        }
        default: return VFVM_ERR_UNREGISTERED;
    }
    switch (ph.slot[VFVM_SLOT_EDGEREACTION].id) {
        case VFVM_NONE:
        case VFVM_EDGEREACTION_DIAMOND: break;  // does not depend on u: no Jacobian entries
        case VFVM_EDGEREACTION_JOULE: {
            const double* p = P(VFVM_SLOT_EDGEREACTION);
            if (p[0] != 0.0) mask_set(m.edgereaction, (int)p[2] * n + (int)p[1]);  // d f_iT / d u_iphi
            break;
        }
        default: return VFVM_ERR_UNREGISTERED;
    }
    for (int e = 0; e < ph.nbc; e++) {
        const vfvm_bc_entry& b = ph.bc[e];
        if (b.kind == VFVM_BC_DIRICHLET || (b.kind == VFVM_BC_ROBIN && b.factor != 0.0)) mask_set(m.boundary, b.species * n + b.species);
    }
    if (ph.has_legacy_bc)
        for (int r = 0; r < ph.nbregions; r++)
            for (int i = 0; i < n; i++)
                if (ph.bfactors[r * n + i] != 0.0) mask_set(m.boundary, i * n + i);
    return VFVM_OK;
}

// extra diagonal-block bits a boundary node in boundary region `breg` receives (for the exported scalar pattern)
static void boundary_bits(const vfvm_handle* h, int breg, uint64_t* out) {
    const int n = h->n;
    const PhysicsDev& ph = h->phys;
    if (ph.slot[VFVM_SLOT_BREACTION].id == VFVM_BREACTION_LINEAR) {
        const double* p = ph.params + ph.slot[VFVM_SLOT_BREACTION].off;
        if ((int)p[0] == breg)
            for (int i = 0; i < n * n; i++)
                if (p[1 + i] != 0.0) mask_set(out, i);
    }
    if (ph.slot[VFVM_SLOT_BREACTION].id == VFVM_BREACTION_POW) {
        const double* p = ph.params + ph.slot[VFVM_SLOT_BREACTION].off;
        if ((int)p[0] == breg)
            for (int i = 0; i < n; i++)
                if (p[1 + i] != 0.0) mask_set(out, i * n + i);
    }
    if (ph.slot[VFVM_SLOT_BREACTION].id == VFVM_BREACTION_CATALYSIS) {
        const double* p = ph.params + ph.slot[VFVM_SLOT_BREACTION].off;
        if ((int)p[0] == breg) {
            const int iA = (int)p[6], iB = (int)p[7], iC = (int)p[8];
            const int pairs[7][2] = {{iA, iA}, {iA, iC}, {iB, iB}, {iB, iC}, {iC, iA}, {iC, iB}, {iC, iC}};
            for (auto& q : pairs) mask_set(out, q[0] * n + q[1]);
        }
    }
    if (ph.slot[VFVM_SLOT_BSTORAGE].id == VFVM_BSTORAGE_LINEAR && h->seen_transient) {
        const double* p = ph.params + ph.slot[VFVM_SLOT_BSTORAGE].off;
        if ((int)p[0] == breg)
            for (int i = 0; i < n; i++)
                if (p[1 + i] != 0.0) mask_set(out, i * n + i);
    }
    for (int e = 0; e < ph.nbc; e++) {
        const vfvm_bc_entry& b = ph.bc[e];
        if (b.region != 0 && b.region != breg) continue;
        if (b.kind == VFVM_BC_DIRICHLET || (b.kind == VFVM_BC_ROBIN && b.factor != 0.0)) mask_set(out, b.species * n + b.species);
    }
    if (ph.has_legacy_bc && breg >= 1 && breg <= ph.nbregions)
        for (int i = 0; i < n; i++)
            if (ph.bfactors[(breg - 1) * n + i] != 0.0) mask_set(out, i * n + i);
}

int vfvm_pattern_build(vfvm_handle* h) {
    const int64_t E = h->E, Nown = h->Nown;
    const int n = h->n;
    const int B = 256;
    cudaStream_t s = h->stream;
    int rc = vfvm_physics_masks(h);
    if (rc) return vfvm_fail(h, rc, "physics id is not in the registered device library");

    // ---- planes
    h->cF = h->cD = 0;
    for (int b = 0; b < n * n; b++) {
        h->idxF[b] = h->idxD[b] = -1;
        if (mask_get(h->masks.flux, b) || mask_get(h->masks.edgereaction, b)) {
            h->idxF[b] = h->cF;
            h->planeF[h->cF++] = b;
        }
        const bool diag = (b / n) == (b % n);
        if (diag || mask_get(h->masks.flux, b) || mask_get(h->masks.edgereaction, b) || mask_get(h->masks.reaction, b) || mask_get(h->masks.storage, b) || mask_get(h->masks.boundary, b)) {
            h->idxD[b] = h->cD;
            h->planeD[h->cD++] = b;
        }
    }

    // ---- directed node graph
    DevBuf<uint64_t> keys;
    DevBuf<int32_t> vals;
    keys.tally = vals.tally = &h->bytes;
    keys.alloc(2 * E);
    vals.alloc(2 * E);
    k_dir_keys<<<cdiv(E, B), B, 0, s>>>(E, Nown, h->edgenodes.p, keys.p, vals.p);
    h->launches++;
    {
        DevBuf<uint64_t> keys2;
        DevBuf<int32_t> vals2;
        keys2.tally = vals2.tally = &h->bytes;
        keys2.alloc(2 * E);
        vals2.alloc(2 * E);
        cub::DoubleBuffer<uint64_t> dk(keys.p, keys2.p);
        cub::DoubleBuffer<int32_t> dv(vals.p, vals2.p);
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, 2 * E, 0, 64, s));
        DevBuf<char> t;
        t.alloc(tmp);
        CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, dk, dv, 2 * E, 0, 64, s));
        CK(cudaStreamSynchronize(s));
        if (dk.Current() != keys.p) std::swap(keys.p, keys2.p);
        if (dv.Current() != vals.p) std::swap(vals.p, vals2.p);
    }
    if (2 * E >= ((int64_t)1 << 31)) throw std::string("pattern too large for 32-bit block indices");
    h->rowptr.alloc(Nown + 1);
    k_rowptr<<<cdiv(Nown + 1, B), B, 0, s>>>(Nown, 2 * E, keys.p, h->rowptr.p);  // keys of halo rows sort behind row Nown-1
    h->launches++;
    {
        int32_t last = 0;
        CK(cudaMemcpyAsync(&last, h->rowptr.p + Nown, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        h->nnz_off = last;
    }
    // ---- SELL-32 layout of the off-diagonal blocks
    const int nslices = (int)((Nown + 31) / 32);
    h->ngroups = nslices;
    {
        DevBuf<int64_t> sp64;
        sp64.alloc((size_t)nslices + 1);
        k_slice_width<<<cdiv(nslices, 8), 256, 0, s>>>(nslices, Nown, h->rowptr.p, sp64.p);
        h->launches++;
        size_t tmp = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, sp64.p, sp64.p, nslices + 1, s));
        DevBuf<char> t;
        t.alloc(tmp);
        CK(cub::DeviceScan::InclusiveSum(t.p, tmp, sp64.p, sp64.p, nslices + 1, s));
        int64_t total = 0;
        CK(cudaMemcpyAsync(&total, sp64.p + nslices, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (total >= ((int64_t)1 << 31)) throw std::string("pattern too large for 32-bit block indices");
        h->nnz_sell = total;
        h->sell_ptr.alloc((size_t)nslices + 1);
        h->colidx.alloc(total);
        h->nz_edge.alloc(total);
        h->nzfac.alloc(total);
        h->upos.alloc(Nown);
        k_fill_sell<<<cdiv(nslices, 8), 256, 0, s>>>(nslices, Nown, h->rowptr.p, sp64.p, keys.p, vals.p, h->ef_colptr.p, h->ef_fac.p, h->sell_ptr.p,
                                                     h->colidx.p, h->nz_edge.p, h->nzfac.p, h->upos.p);
        h->launches++;
        CK(cudaStreamSynchronize(s));
        std::vector<int32_t> sl = h->sell_ptr.to_host(s);  // widest slice = longest row: the row kernels pick their batch size by it
        h->group_maxnnz = 0;
        for (int g = 0; g < nslices; g++) h->group_maxnnz = std::max(h->group_maxnnz, (sl[g + 1] - sl[g]) / 32);
    }
    keys.release();
    vals.release();

    // ---- boundary nodes: node -> (bface, local node) in ascending bface order (the reference's loop order)
    {
        std::vector<int32_t> bfn = h->bfacenodes.to_host(s), bfr = h->bfaceregions.to_host(s);
        const int dim = h->dim;
        std::vector<std::pair<int32_t, int32_t>> items;
        items.reserve(bfn.size());
        for (int64_t i = 0; i < (int64_t)bfn.size(); i++)
            if (bfn[i] < Nown) items.emplace_back(bfn[i], (int32_t)i);
        std::stable_sort(items.begin(), items.end(), [](auto& a, auto& b) { return a.first < b.first; });
        std::vector<int32_t> node, ptr, bface, local;
        std::vector<uint8_t> bm;
        for (size_t i = 0; i < items.size(); i++) {
            if (i == 0 || items[i].first != items[i - 1].first) {
                node.push_back(items[i].first);
                ptr.push_back((int32_t)i);
            }
            bface.push_back(items[i].second / dim);
            local.push_back(items[i].second % dim);
        }
        ptr.push_back((int32_t)items.size());
        h->nbnodes = (int64_t)node.size();
        h->nbitems = (int64_t)items.size();
        if (h->masked) {  // species defined per node = union over the cell regions the node touches (node_dof, src/vfvm_system.jl:445-456)
            std::vector<int64_t> cp = h->nf_colptr.to_host(s);
            std::vector<int32_t> rg = h->nf_region.to_host(s), act((size_t)h->N, 0);
            for (int64_t K = 0; K < h->N; K++)
                for (int64_t q = cp[K]; q < cp[K + 1]; q++)
                    for (int i = 0; i < h->n; i++)
                        if (h->region_species[(size_t)(rg[q] - 1) * h->n + i]) act[K] |= (1 << i);
            if (!h->bregion_species.empty())  // boundary species are defined at the nodes of their boundary regions (src/vfvm_system.jl:502-513)
                for (int64_t b = 0; b < (int64_t)bfr.size(); b++) {
                    const int br = bfr[b] - 1;
                    if (br >= h->nbregions_bs) continue;
                    for (int i = 0; i < h->n; i++)
                        if (h->bregion_species[(size_t)br * h->n + i])
                            for (int l = 0; l < dim; l++) act[bfn[b * dim + l]] |= (1 << i);
                }
            h->node_active_host = act;
            h->node_active.upload(act.data(), act.size(), s);
        } else {
            h->node_active.release();
            h->node_active_host.clear();
        }
        h->bn_node_host = node;
        h->pipe.valid = false;
        h->bn_node.upload(node.data(), node.size(), s);
        h->bn_ptr.upload(ptr.data(), ptr.size(), s);
        h->bn_bface.upload(bface.data(), bface.size(), s);
        h->bn_local.upload(local.data(), local.size(), s);
        // per boundary node: diag bits contributed by boundary terms
        h->bnode_mask_host.assign((size_t)h->nbnodes * 16, 0);
        for (int64_t b = 0; b < h->nbnodes; b++) {
            uint64_t bits[2] = {0, 0};
            for (int32_t q = ptr[b]; q < ptr[b + 1]; q++) boundary_bits(h, bfr[bface[q]], bits);
            memcpy(&h->bnode_mask_host[(size_t)b * 16], bits, 16);
        }
        CK(cudaStreamSynchronize(s));
    }

    // ---- values + vectors
    h->offval.alloc((size_t)std::max(1, h->cF) * h->nnz_sell);
    h->diagval.alloc((size_t)h->cD * Nown);
    for (int v = 0; v < 4; v++) {
        h->vec[v].alloc((size_t)n * h->N);
        CK(cudaMemsetAsync(h->vec[v].p, 0, sizeof(double) * n * h->N, s));
    }
    CK(cudaMemsetAsync(h->offval.p, 0, sizeof(double) * h->offval.n, s));
    CK(cudaMemsetAsync(h->diagval.p, 0, sizeof(double) * h->diagval.n, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    h->seen_transient = false;
    h->precon_valid = false;
    h->ilu_struct_valid = false;
    h->graph_epoch++;
    h->have_pattern = true;
    return VFVM_OK;
}

// ---- scalar pattern / value export (host side; parity + interop, not on the hot path) -----------------------------
struct ScalarPattern {
    std::vector<int64_t> rowptr, colidx;
    std::vector<int64_t> src;  // >= 0: offval index (plane*nnz_off + k) ; < 0: -(1 + diag index (plane*Nown + K))
};

// scalar CSR rows of the nodes [K0, K1); only the slices / planes of that range are downloaded (a probe of a few node planes of a
// 193^3 problem must not pull the whole Jacobian over PCIe)
template <class T>
static std::vector<T> fetch_range(const DevBuf<T>& b, int64_t i0, int64_t i1, cudaStream_t s) {
    std::vector<T> v((size_t)std::max<int64_t>(0, i1 - i0));
    if (!v.empty()) CK(cudaMemcpyAsync(v.data(), b.p + i0, v.size() * sizeof(T), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return v;
}

static void build_scalar_range(vfvm_handle* h, int64_t K0, int64_t K1, ScalarPattern& sp, int64_t* e0_out = nullptr, int64_t* e1_out = nullptr) {
    const int n = h->n;
    const int64_t Nown = h->Nown;
    const int64_t g0 = K0 >> 5, g1 = (K1 + 31) >> 5;
    std::vector<int32_t> rp = fetch_range(h->rowptr, K0, K1 + 1, h->stream), sl = fetch_range(h->sell_ptr, g0, g1 + 1, h->stream), bn = h->bn_node.to_host(h->stream);
    const int64_t e0 = sl.empty() ? 0 : sl.front(), e1 = sl.empty() ? 0 : sl.back();
    if (e0_out) *e0_out = e0;
    if (e1_out) *e1_out = e1;
    std::vector<int32_t> ci = fetch_range(h->colidx, e0, e1, h->stream);
    std::vector<int64_t> bnode_of((size_t)(K1 - K0), -1);
    for (size_t b = 0; b < bn.size(); b++)
        if (bn[b] >= K0 && bn[b] < K1) bnode_of[bn[b] - K0] = (int64_t)b;
    // masked systems: species pairs that share a cell region at the node / on the edge (the reference never touches other entries)
    std::vector<int64_t> nfp, efp;
    std::vector<int32_t> nfr, efr, nze;
    if (h->masked) {
        nfp = h->nf_colptr.to_host(h->stream);
        nfr = h->nf_region.to_host(h->stream);
        efp = h->ef_colptr.to_host(h->stream);
        efr = h->ef_region.to_host(h->stream);
        nze = fetch_range(h->nz_edge, e0, e1, h->stream);
    }
    auto pair_ok = [&](const std::vector<int64_t>& ptr, const std::vector<int32_t>& reg, int64_t item, int i, int j) {
        if (!h->masked) return true;
        for (int64_t q = ptr[item]; q < ptr[item + 1]; q++)
            if (h->region_species[(size_t)(reg[q] - 1) * n + i] && h->region_species[(size_t)(reg[q] - 1) * n + j]) return true;
        return false;
    };
    sp.rowptr.assign((size_t)(K1 - K0) * n + 1, 0);
    sp.colidx.clear();
    sp.src.clear();
    const uint64_t offm[2] = {h->masks.flux[0] | h->masks.edgereaction[0], h->masks.flux[1] | h->masks.edgereaction[1]};
    for (int64_t K = K0; K < K1; K++) {
        uint64_t dm[2] = {offm[0] | h->masks.reaction[0], offm[1] | h->masks.reaction[1]};  // node and edge terms
        if (h->seen_transient) {
            dm[0] |= h->masks.storage[0];
            dm[1] |= h->masks.storage[1];
        }
        uint64_t bm[2] = {0, 0};  // boundary terms of this node
        if (bnode_of[K - K0] >= 0) memcpy(bm, &h->bnode_mask_host[(size_t)bnode_of[K - K0] * 16], 16);
        // species defined at the node: every species, or (masked systems) the union over its cell regions and boundary regions
        const unsigned act = h->masked ? (unsigned)h->node_active_host[(size_t)K] : 0xffffffffu;
        const int len = rp[K - K0 + 1] - rp[K - K0];
        const int64_t ebase = (int64_t)sl[(K >> 5) - g0] + (K & 31);
        for (int i = 0; i < n; i++) {
            bool diag_done = false;
            const bool inactive = !((act >> i) & 1u);  // identity row
            auto emit_diag = [&]() {
                for (int j = 0; j < n; j++) {
                    const bool node_term = mask_get(dm, i * n + j) && pair_ok(nfp, nfr, K, i, j);
                    const bool bnd_term = mask_get(bm, i * n + j) && ((act >> j) & 1u);  // assemble_res_jac(bnode): both species defined at the node
                    if (inactive ? (i == j) : (node_term || bnd_term)) {
                        sp.colidx.push_back(K * n + j);
                        sp.src.push_back(-(1 + (int64_t)h->idxD[i * n + j] * Nown + K));
                    }
                }
                diag_done = true;
            };
            for (int q = 0; q < len; q++) {
                const int64_t e = ebase + (int64_t)q * 32;
                const int64_t L = ci[e - e0];
                if (!diag_done && L > K) emit_diag();
                for (int j = 0; j < n; j++)
                    if (mask_get(offm, i * n + j) && (!h->masked || (nze[e - e0] >= 0 && pair_ok(efp, efr, nze[e - e0], i, j)))) {
                        sp.colidx.push_back(L * n + j);
                        sp.src.push_back((int64_t)h->idxF[i * n + j] * h->nnz_sell + e);
                    }
            }
            if (!diag_done) emit_diag();
            sp.rowptr[(K - K0) * n + i + 1] = (int64_t)sp.colidx.size();
        }
    }
}

static void build_scalar(vfvm_handle* h, ScalarPattern& sp) { build_scalar_range(h, 0, h->Nown, sp); }

static int need_pattern(vfvm_handle* h) {
    if (!h || !h->have_pattern) return vfvm_fail(h, VFVM_ERR_STATE, "vfvm_build_pattern has not been called");
    return 0;
}

extern "C" int vfvm_pattern_size(vfvm_handle* h, int64_t* nrows, int64_t* nnz) {
    if (int rc = need_pattern(h)) return rc;
    VFVM_TRY(h, {
        ScalarPattern sp;
        build_scalar(h, sp);
        *nrows = h->Nown * h->n;
        *nnz = (int64_t)sp.colidx.size();
    })
    return VFVM_OK;
}

extern "C" int vfvm_get_pattern_csr(vfvm_handle* h, int64_t* rowptr, int64_t* colidx) {
    if (int rc = need_pattern(h)) return rc;
    VFVM_TRY(h, {
        ScalarPattern sp;
        build_scalar(h, sp);
        std::copy(sp.rowptr.begin(), sp.rowptr.end(), rowptr);
        std::copy(sp.colidx.begin(), sp.colidx.end(), colidx);
    })
    return VFVM_OK;
}

// CSR -> CSC permutation: perm[c] = CSR position of CSC entry c
static void csr_to_csc(const ScalarPattern& sp, int64_t ncols, std::vector<int64_t>& colptr, std::vector<int64_t>& rowval, std::vector<int64_t>& perm) {
    const int64_t nrows = (int64_t)sp.rowptr.size() - 1, nnz = (int64_t)sp.colidx.size();
    colptr.assign((size_t)ncols + 1, 0);
    for (int64_t k = 0; k < nnz; k++) colptr[sp.colidx[k] + 1]++;
    for (int64_t c = 0; c < ncols; c++) colptr[c + 1] += colptr[c];
    rowval.resize(nnz);
    perm.resize(nnz);
    std::vector<int64_t> fill(colptr.begin(), colptr.end() - 1);
    for (int64_t r = 0; r < nrows; r++)
        for (int64_t k = sp.rowptr[r]; k < sp.rowptr[r + 1]; k++) {
            const int64_t o = fill[sp.colidx[k]]++;
            rowval[o] = r;
            perm[o] = k;
        }
}

extern "C" int vfvm_get_pattern_csc(vfvm_handle* h, int64_t* colptr, int64_t* rowval) {
    if (int rc = need_pattern(h)) return rc;
    VFVM_TRY(h, {
        ScalarPattern sp;
        build_scalar(h, sp);
        std::vector<int64_t> cp, rv, perm;
        csr_to_csc(sp, h->N * h->n, cp, rv, perm);
        std::copy(cp.begin(), cp.end(), colptr);
        std::copy(rv.begin(), rv.end(), rowval);
    })
    return VFVM_OK;
}

static int get_nzval(vfvm_handle* h, double* out, int memspace, bool csc) {
    if (int rc = need_pattern(h)) return rc;
    VFVM_TRY(h, {
        ScalarPattern sp;
        build_scalar(h, sp);
        std::vector<double> off = h->offval.to_host(h->stream), dg = h->diagval.to_host(h->stream);
        std::vector<double> v(sp.src.size());
        for (size_t k = 0; k < sp.src.size(); k++) v[k] = sp.src[k] >= 0 ? off[sp.src[k]] : dg[-(sp.src[k] + 1)];
        if (csc) {
            std::vector<int64_t> cp, rv, perm;
            csr_to_csc(sp, h->N * h->n, cp, rv, perm);
            std::vector<double> w(v.size());
            for (size_t c = 0; c < v.size(); c++) w[c] = v[perm[c]];
            v.swap(w);
        }
        if (memspace == VFVM_HOST) std::copy(v.begin(), v.end(), out);
        else CK(cudaMemcpy(out, v.data(), v.size() * 8, cudaMemcpyHostToDevice));
    })
    return VFVM_OK;
}

extern "C" int vfvm_get_nzval_csr(vfvm_handle* h, double* nzval, int memspace) { return get_nzval(h, nzval, memspace, false); }
extern "C" int vfvm_get_nzval_csc(vfvm_handle* h, double* nzval, int memspace) { return get_nzval(h, nzval, memspace, true); }

// Scalar CSR rows of the owned nodes [node0, node1) with their values (parity probes at full problem size compare a few node
// planes of the 193^3 Jacobians entry by entry with a CPU evaluation).  Call with colidx == NULL to get the entry count first.
// rowptr: (node1 - node0) * n + 1 offsets starting at 0; colidx: local dof numbers.
extern "C" int vfvm_get_rows_csr(vfvm_handle* h, int64_t node0, int64_t node1, int64_t* nnz_out, int64_t* rowptr, int64_t* colidx, double* nzval) {
    if (int rc = need_pattern(h)) return rc;
    if (node0 < 0 || node1 > h->Nown || node0 >= node1 || !nnz_out) return vfvm_fail(h, VFVM_ERR_ARG, "bad node range");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        ScalarPattern sp;
        int64_t e0 = 0, e1 = 0;
        build_scalar_range(h, node0, node1, sp, &e0, &e1);
        *nnz_out = (int64_t)sp.colidx.size();
        if (!colidx) return VFVM_OK;
        std::copy(sp.rowptr.begin(), sp.rowptr.end(), rowptr);
        std::copy(sp.colidx.begin(), sp.colidx.end(), colidx);
        std::vector<std::vector<double>> off((size_t)h->cF), dg((size_t)h->cD);
        for (int p = 0; p < h->cF; p++) off[p] = fetch_range(h->offval, (int64_t)p * h->nnz_sell + e0, (int64_t)p * h->nnz_sell + e1, h->stream);
        for (int p = 0; p < h->cD; p++) dg[p] = fetch_range(h->diagval, (int64_t)p * h->Nown + node0, (int64_t)p * h->Nown + node1, h->stream);
        for (size_t k = 0; k < sp.src.size(); k++) {
            if (sp.src[k] >= 0) {
                const int64_t p = sp.src[k] / h->nnz_sell, e = sp.src[k] - p * h->nnz_sell;
                nzval[k] = off[(size_t)p][(size_t)(e - e0)];
            } else {
                const int64_t q = -(sp.src[k] + 1), p = q / h->Nown, K = q - p * h->Nown;
                nzval[k] = dg[(size_t)p][(size_t)(K - node0)];
            }
        }
    })
    return VFVM_OK;
}
