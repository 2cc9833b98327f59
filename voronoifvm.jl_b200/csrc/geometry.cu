// K1 + K2: edge enumeration and Voronoi form factors on the device.
//   K1 replaces ExtendableGrids' CellEdges/EdgeNodes instantiation triggered at src/vfvm_system.jl:702-703.
//   K2 replaces cellfactors!/bfacefactors! (src/vfvm_formfactors.jl:12-332) and the accumulation loop of
//      update_grid_edgewise! (src/vfvm_system.jl:716-735).
// Compiled with -fmad=false so that the per-simplex formulas round exactly like a plain fp64 CPU evaluation.
// Accumulation over the cells around an edge / node is a segmented sum in ascending cell order (the reference's
// loop order), obtained from a stable radix sort -- deterministic, no atomics.
// CUB is used for the one-off sorts/scans; every kernel here is ours.
#include <cub/cub.cuh>

#include "vfvm_internal.h"

namespace {

__device__ __constant__ int c_len2[3][2] = {{1, 2}, {2, 0}, {0, 1}};
__device__ __constant__ int c_len3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

__device__ __forceinline__ void local_edge(int dim, int ie, int& a, int& b) {
    if (dim == 1) {
        a = 0;
        b = 1;
    } else if (dim == 2) {
        a = c_len2[ie][0];
        b = c_len2[ie][1];
    } else {
        a = c_len3[ie][0];
        b = c_len3[ie][1];
    }
}

// ---- K1 ------------------------------------------------------------------------------------------------
__global__ void k_edge_keys(int dim, int64_t C, const int32_t* __restrict__ cellnodes, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int nn = dim + 1, ne = dim * (dim + 1) / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * ne) return;
    const int64_t c = i / ne;
    const int ie = (int)(i - c * ne);
    int a, b;
    local_edge(dim, ie, a, b);
    const uint32_t na = cellnodes[c * nn + a], nb = cellnodes[c * nn + b];
    const uint64_t lo = min(na, nb), hi = max(na, nb);
    keys[i] = (lo << 32) | hi;  // CSC order of the lower triangle: by smaller node, then larger node
    vals[i] = (int32_t)i;
}

__global__ void k_head_flags(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// rank[i] = inclusive scan of flags; edge id = rank-1
__global__ void k_edges_from_sorted(int64_t n, const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals, const int32_t* __restrict__ rank,
                                    int32_t* __restrict__ celledges, int32_t* __restrict__ edgenodes, int64_t* __restrict__ seg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t e = rank[i] - 1;
    celledges[vals[i]] = e;
    if (i == 0 || keys[i] != keys[i - 1]) {
        edgenodes[2 * (int64_t)e] = (int32_t)(keys[i] & 0xffffffffu);  // edge.node[1]: larger node
        edgenodes[2 * (int64_t)e + 1] = (int32_t)(keys[i] >> 32);      // edge.node[2]: smaller node
        seg[e] = i;
    }
    if (i == n - 1) seg[e + 1] = n;
}

__global__ void k_node_keys(int dim, int64_t C, const int32_t* __restrict__ cellnodes, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int nn = dim + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * nn) return;
    keys[i] = (uint32_t)cellnodes[i];
    vals[i] = (int32_t)i;
}

__global__ void k_lower_bound(int64_t nitems, int64_t nkeys, const uint32_t* __restrict__ keys, int64_t* __restrict__ seg) {
    const int64_t K = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (K > nitems) return;
    int64_t lo = 0, hi = nkeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys[mid] < K) lo = mid + 1;
        else hi = mid;
    }
    seg[K] = lo;
}

// ---- K2: per-simplex factors ----------------------------------------------------------------------------
__device__ void cellfactors_dev(int dim, int coordsys, const double* __restrict__ coord, const int32_t* __restrict__ n, double* npar, double* epar) {
    const double PI = 3.14159265358979323846;
    if (dim == 1) {
        const double xK = coord[n[0]], xL = coord[n[1]];
        if (coordsys == VFVM_CARTESIAN) {
            const double d = fabs(xL - xK);
            npar[0] = d / 2;
            npar[1] = d / 2;
            epar[0] = 1 / d;
        } else {
            double r0 = xK, r1 = xL;
            if (r1 < r0) {
                r0 = xL;
                r1 = xK;
            }
            const double rhalf = 0.5 * (r1 + r0);
            if (coordsys == VFVM_CYLINDRICAL) {
                npar[0] = PI * (rhalf * rhalf - r0 * r0);
                npar[1] = PI * (r1 * r1 - rhalf * rhalf);
                epar[0] = 2.0 * PI * rhalf / (r1 - r0);
            } else {
                npar[0] = PI * (rhalf * rhalf * rhalf - r0 * r0 * r0) * 4.0 / 3.0;
                npar[1] = PI * (r1 * r1 * r1 - rhalf * rhalf * rhalf) * 4.0 / 3.0;
                epar[0] = 4.0 * PI * (rhalf * rhalf) / (r1 - r0);
            }
        }
        return;
    }
    if (dim == 2) {
        double V[2][3], dd[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int a = n[c_len2[i][0]], b = n[c_len2[i][1]];
            V[0][i] = coord[2 * (int64_t)a] - coord[2 * (int64_t)b];
            V[1][i] = coord[2 * (int64_t)a + 1] - coord[2 * (int64_t)b + 1];
            dd[i] = V[0][i] * V[0][i] + V[1][i] * V[1][i];
        }
        const double det = V[0][2] * V[1][1] - V[0][1] * V[1][2];
        const double vol = fabs(0.5 * det);
        const double ivol = 1.0 / vol;
        epar[0] = (dd[1] + dd[2] - dd[0]) * 0.125 * ivol;
        epar[1] = (dd[2] + dd[0] - dd[1]) * 0.125 * ivol;
        epar[2] = (dd[0] + dd[1] - dd[2]) * 0.125 * ivol;
        npar[0] = npar[1] = npar[2] = 0.0;
        if (coordsys == VFVM_CARTESIAN) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                npar[c_len2[i][0]] += epar[i] * dd[i] * 0.25;
                npar[c_len2[i][1]] += epar[i] * dd[i] * 0.25;
            }
        } else {
            // circumcentre radius (ExtendableGrids.tricircumcenter!, Shewchuk)
            const double* A = &coord[2 * (int64_t)n[0]];
            const double* B = &coord[2 * (int64_t)n[1]];
            const double* Cc = &coord[2 * (int64_t)n[2]];
            const double xba = B[0] - A[0], yba = B[1] - A[1], xca = Cc[0] - A[0], yca = Cc[1] - A[1];
            const double balength = xba * xba + yba * yba, calength = xca * xca + yca * yca;
            const double denominator = 0.5 / (xba * yca - yba * xca);
            const double rcc = (yca * balength - yba * calength) * denominator + A[0];
            double emid[3];
#pragma unroll
            for (int i = 0; i < 3; i++) emid[i] = 0.5 * (coord[2 * (int64_t)n[c_len2[i][0]]] + coord[2 * (int64_t)n[c_len2[i][1]]]);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double r1 = coord[2 * (int64_t)n[c_len2[i][0]]], r2 = coord[2 * (int64_t)n[c_len2[i][1]]];
                const double cylfac1 = 2 * PI * (r1 + rcc + emid[i]) / 3;
                const double cylfac2 = 2 * PI * (r2 + rcc + emid[i]) / 3;
                npar[c_len2[i][0]] += epar[i] * dd[i] * 0.25 * cylfac1;
                npar[c_len2[i][1]] += epar[i] * dd[i] * 0.25 * cylfac2;
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double rmid = (rcc + emid[i]) / 2;
                epar[i] *= 2 * PI * rmid;
            }
        }
        return;
    }
    // Tetrahedron3D, src/vfvm_formfactors.jl:163-235
    const int pi1[4] = {4, 5, 4, 0}, pi2[4] = {5, 2, 0, 3}, pi3[4] = {3, 1, 2, 1};
    const int po1[4] = {1, 0, 1, 5}, po2[4] = {0, 3, 5, 2}, po3[4] = {2, 4, 3, 4};
    double X[4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int d = 0; d < 3; d++) X[i][d] = coord[3 * (int64_t)n[i] + d];
    double dd[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const int a = c_len3[i][0], b = c_len3[i][1];
        const double dx = X[a][0] - X[b][0], dy = X[a][1] - X[b][1], dz = X[a][2] - X[b][2];
        dd[i] = dx * dx + dy * dy + dz * dz;
        epar[i] = 0.0;
    }
    const double x1 = X[1][0] - X[0][0], y1 = X[1][1] - X[0][1], z1 = X[1][2] - X[0][2];
    const double x2 = X[2][0] - X[0][0], y2 = X[2][1] - X[0][1], z2 = X[2][2] - X[0][2];
    const double x3 = X[3][0] - X[0][0], y3 = X[3][1] - X[0][1], z3 = X[3][2] - X[0][2];
    double det = (x1 * (y2 * z3 - y3 * z2) + x2 * (y3 * z1 - y1 * z3) + x3 * (y1 * z2 - y2 * z1));
    if (det < 0) det = -det;
    const double vol = det / 6;
    const double vv = 96 * 6 * vol;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        npar[i] = 0.0;
        const int i1 = pi1[i], i2 = pi2[i], i3 = pi3[i];
        const double h1 = dd[i1] * (dd[i2] + dd[i3] - dd[i1]);
        const double h2 = dd[i2] * (dd[i3] + dd[i1] - dd[i2]);
        const double h3 = dd[i3] * (dd[i1] + dd[i2] - dd[i3]);
        const double df = h1 + h2 + h3;
        const double vf = (h1 * dd[po1[i]] + h2 * dd[po2[i]] + h3 * dd[po3[i]] - 2 * dd[i1] * dd[i2] * dd[i3]) / (vv * df);
        epar[i1] += h1 * vf;
        epar[i2] += h2 * vf;
        epar[i3] += h3 * vf;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        npar[c_len3[i][0]] += epar[i];
        npar[c_len3[i][1]] += epar[i];
        epar[i] = 6 * epar[i] / dd[i];
    }
}

__global__ void k_cellfactors(int dim, int coordsys, int64_t C, const double* __restrict__ coord, const int32_t* __restrict__ cellnodes,
                              double* __restrict__ cell_npar, double* __restrict__ cell_epar) {
    const int nn = dim + 1, ne = dim * (dim + 1) / 2;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double npar[4], epar[6];
    int32_t n[4];
    for (int i = 0; i < nn; i++) n[i] = cellnodes[c * nn + i];
    cellfactors_dev(dim, coordsys, coord, n, npar, epar);
    for (int i = 0; i < nn; i++) cell_npar[c * nn + i] = npar[i];
    for (int i = 0; i < ne; i++) cell_epar[c * ne + i] = epar[i];
}

__global__ void k_bfacefactors(int dim, int coordsys, int64_t NB, const double* __restrict__ coord, const int32_t* __restrict__ bfacenodes,
                               double* __restrict__ bfnf) {
    const double PI = 3.14159265358979323846;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    const int32_t* n = &bfacenodes[b * dim];
    if (dim == 1) {
        const double r = coord[n[0]];
        bfnf[b] = coordsys == VFVM_CARTESIAN ? 1.0 : (coordsys == VFVM_CYLINDRICAL ? 2 * PI * r : 4 * PI * (r * r));
    } else if (dim == 2) {
        const int i1 = n[0], i2 = n[1];
        if (coordsys == VFVM_CARTESIAN) {
            const double dx = coord[2 * (int64_t)i1] - coord[2 * (int64_t)i2], dy = coord[2 * (int64_t)i1 + 1] - coord[2 * (int64_t)i2 + 1];
            const double d = sqrt(dx * dx + dy * dy);
            bfnf[2 * b] = d / 2;
            bfnf[2 * b + 1] = d / 2;
        } else {
            const double r1 = coord[2 * (int64_t)i1], r2 = coord[2 * (int64_t)i2], z1 = coord[2 * (int64_t)i1 + 1], z2 = coord[2 * (int64_t)i2 + 1];
            const double dr = r1 - r2, rmid = (r1 + r2) / 2, dz = z1 - z2;
            const double l = sqrt(dr * dr + dz * dz);
            bfnf[2 * b] = PI * (r1 + rmid) * l / 2;
            bfnf[2 * b + 1] = PI * (r2 + rmid) * l / 2;
        }
    } else {
        double epar[3] = {0, 0, 0}, npar[3] = {0, 0, 0};
        for (int j = 0; j < 3; j++) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double d = coord[3 * (int64_t)n[c_len2[i][0]] + j] - coord[3 * (int64_t)n[c_len2[i][1]] + j];
                epar[i] += d * d;
            }
        }
        const double dd[3] = {epar[0], epar[1], epar[2]};
        epar[0] = (dd[1] + dd[2] - dd[0]) * dd[0];
        epar[1] = (dd[2] + dd[0] - dd[1]) * dd[1];
        epar[2] = (dd[0] + dd[1] - dd[2]) * dd[2];
        const double vol = sqrt(epar[0] + epar[1] + epar[2]) * 0.25;
        const double d = 1.0 / (8 * vol);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            npar[c_len2[i][0]] += epar[i] * d * 0.25;
            npar[c_len2[i][1]] += epar[i] * d * 0.25;
        }
        for (int i = 0; i < 3; i++) bfnf[3 * b + i] = npar[i];
    }
}

// ---- K2: segmented accumulation into the nregions x nitems CSC factor matrices --------------------------
#define MAXR 8
// pass 0: count distinct regions per item; pass 1: fill (region ascending, sum in ascending cell order)
__global__ void k_accumulate(int pass, int per_cell, int64_t nitems, const int64_t* __restrict__ seg, const int32_t* __restrict__ vals,
                             const int32_t* __restrict__ cellregions, const double* __restrict__ cellfac, int64_t* __restrict__ colptr,
                             int32_t* __restrict__ region, double* __restrict__ fac, int32_t* __restrict__ errflag) {
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= nitems) return;
    int regs[MAXR];
    double sums[MAXR];
    int nr = 0;
    for (int64_t j = seg[it]; j < seg[it + 1]; j++) {
        const int32_t idx = vals[j];
        const int r = cellregions[idx / per_cell];
        int q = 0;
        while (q < nr && regs[q] != r) q++;
        if (q == nr) {
            if (nr == MAXR) {
                *errflag = 1;
                continue;
            }
            regs[nr] = r;
            sums[nr] = 0.0;
            nr++;
        }
        if (pass == 1) sums[q] += cellfac[idx];
    }
    if (pass == 0) {
        colptr[it + 1] = nr;  // turned into offsets by an inclusive scan
        if (it == 0) colptr[0] = 0;
        return;
    }
    // insertion sort by region label
    for (int a = 1; a < nr; a++) {
        const int r = regs[a];
        const double s = sums[a];
        int b = a - 1;
        while (b >= 0 && regs[b] > r) {
            regs[b + 1] = regs[b];
            sums[b + 1] = sums[b];
            b--;
        }
        regs[b + 1] = r;
        sums[b + 1] = s;
    }
    const int64_t o = colptr[it];
    for (int q = 0; q < nr; q++) {
        region[o + q] = regs[q];
        fac[o + q] = sums[q];
    }
}

template <class K>
void sort_pairs(vfvm_handle* h, int64_t n, DevBuf<K>& keys, DevBuf<int32_t>& vals, int end_bit) {
    DevBuf<K> keys2;
    DevBuf<int32_t> vals2;
    keys2.tally = vals2.tally = &h->bytes;
    keys2.alloc(n);
    vals2.alloc(n);
    cub::DoubleBuffer<K> dk(keys.p, keys2.p);
    cub::DoubleBuffer<int32_t> dv(vals.p, vals2.p);
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, n, 0, end_bit, h->stream));
    DevBuf<char> t;
    t.tally = &h->bytes;
    t.alloc(tmp);
    CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, dk, dv, n, 0, end_bit, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (dk.Current() != keys.p) {
        std::swap(keys.p, keys2.p);
    }
    if (dv.Current() != vals.p) {
        std::swap(vals.p, vals2.p);
    }
}

void inclusive_scan_i64(vfvm_handle* h, int64_t* p, int64_t n) {
    size_t tmp = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, p, p, n, h->stream));
    DevBuf<char> t;
    t.alloc(tmp);
    CK(cub::DeviceScan::InclusiveSum(t.p, tmp, p, p, n, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

int bits_for(uint64_t maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b)) b++;
    return b;
}

void build_factors(vfvm_handle* h, int per_cell, int64_t nitems, const DevBuf<int64_t>& seg, const DevBuf<int32_t>& vals, const DevBuf<double>& cellfac,
                   DevBuf<int64_t>& colptr, DevBuf<int32_t>& region, DevBuf<double>& fac) {
    DevBuf<int32_t> errflag;
    errflag.alloc(1);
    CK(cudaMemsetAsync(errflag.p, 0, 4, h->stream));
    colptr.alloc(nitems + 1);
    const int B = 256;
    k_accumulate<<<cdiv(nitems, B), B, 0, h->stream>>>(0, per_cell, nitems, seg.p, vals.p, h->cellregions.p, cellfac.p, colptr.p, nullptr, nullptr, errflag.p);
    h->launches++;
    inclusive_scan_i64(h, colptr.p, nitems + 1);
    int64_t total = 0;
    CK(cudaMemcpy(&total, colptr.p + nitems, 8, cudaMemcpyDeviceToHost));
    region.alloc(total);
    fac.alloc(total);
    k_accumulate<<<cdiv(nitems, B), B, 0, h->stream>>>(1, per_cell, nitems, seg.p, vals.p, h->cellregions.p, cellfac.p, colptr.p, region.p, fac.p, errflag.p);
    h->launches++;
    int32_t ef = 0;
    CK(cudaMemcpy(&ef, errflag.p, 4, cudaMemcpyDeviceToHost));
    if (ef) throw std::string("more than 8 cell regions meet in one node/edge");
}

}  // namespace

int vfvm_geometry_build(vfvm_handle* h) {
    const int dim = h->dim, nn = dim + 1, ne = dim * (dim + 1) / 2;
    const int64_t C = h->C, N = h->N;
    const int B = 256;
    cudaStream_t s = h->stream;

    // ---- K1: edges
    const int64_t nk = C * ne;
    if (nk >= (int64_t)1 << 31) throw std::string("grid too large for 32-bit (cell, local edge) ids");
    DevBuf<uint64_t> keys;
    DevBuf<int32_t> vals, rank;
    keys.tally = vals.tally = rank.tally = &h->bytes;
    keys.alloc(nk);
    vals.alloc(nk);
    k_edge_keys<<<cdiv(nk, B), B, 0, s>>>(dim, C, h->cellnodes.p, keys.p, vals.p);
    h->launches++;
    sort_pairs<uint64_t>(h, nk, keys, vals, 32 + bits_for((uint64_t)N));
    rank.alloc(nk);
    k_head_flags<<<cdiv(nk, B), B, 0, s>>>(nk, keys.p, rank.p);
    h->launches++;
    {
        size_t tmp = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, rank.p, rank.p, nk, s));
        DevBuf<char> t;
        t.alloc(tmp);
        CK(cub::DeviceScan::InclusiveSum(t.p, tmp, rank.p, rank.p, nk, s));
    }
    int32_t E32 = 0;
    CK(cudaMemcpyAsync(&E32, rank.p + nk - 1, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h->E = E32;
    h->edgenodes.alloc(2 * h->E);
    h->celledges.alloc(nk);
    DevBuf<int64_t> eseg;
    eseg.tally = &h->bytes;
    eseg.alloc(h->E + 1);
    k_edges_from_sorted<<<cdiv(nk, B), B, 0, s>>>(nk, keys.p, vals.p, rank.p, h->celledges.p, h->edgenodes.p, eseg.p);
    h->launches++;
    keys.release();
    rank.release();

    // ---- K2: per-cell factors, then segmented sums
    DevBuf<double> cell_npar, cell_epar;
    cell_npar.tally = cell_epar.tally = &h->bytes;
    cell_npar.alloc(C * nn);
    cell_epar.alloc(C * ne);
    k_cellfactors<<<cdiv(C, B), B, 0, s>>>(dim, h->coordsys, C, h->coord.p, h->cellnodes.p, cell_npar.p, cell_epar.p);
    h->launches++;
    build_factors(h, ne, h->E, eseg, vals, cell_epar, h->ef_colptr, h->ef_region, h->ef_fac);
    eseg.release();
    vals.release();
    cell_epar.release();

    // node -> cells incidence
    const int64_t nkn = C * nn;
    DevBuf<uint32_t> nkeys;
    DevBuf<int32_t> nvals;
    DevBuf<int64_t> nseg;
    nkeys.tally = nvals.tally = nseg.tally = &h->bytes;
    nkeys.alloc(nkn);
    nvals.alloc(nkn);
    k_node_keys<<<cdiv(nkn, B), B, 0, s>>>(dim, C, h->cellnodes.p, nkeys.p, nvals.p);
    h->launches++;
    sort_pairs<uint32_t>(h, nkn, nkeys, nvals, bits_for((uint64_t)N));
    nseg.alloc(N + 1);
    k_lower_bound<<<cdiv(N + 1, B), B, 0, s>>>(N, nkn, nkeys.p, nseg.p);
    h->launches++;
    build_factors(h, nn, N, nseg, nvals, cell_npar, h->nf_colptr, h->nf_region, h->nf_fac);

    // ---- boundary faces
    h->bfacenodefac.alloc(h->NB * dim);
    if (h->NB) {
        k_bfacefactors<<<cdiv(h->NB, B), B, 0, s>>>(dim, h->coordsys, h->NB, h->coord.p, h->bfacenodes.p, h->bfacenodefac.p);
        h->launches++;
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    h->single_region = (h->ef_fac.n == (size_t)h->E) && (h->nf_fac.n == (size_t)N) && h->ncellregions >= 1;
    // single_region additionally requires that every cell carries the same label
    h->have_geometry = true;
    return VFVM_OK;
}
