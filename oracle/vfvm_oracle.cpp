// TEST INFRASTRUCTURE (oracle) -- never imported, linked or executed by the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// CPU restatement of the Newton hot path of VoronoiFVM.jl v3.5.2 (pure Julia, cannot run in this image: no
// julia binary, un-vendored dependencies).  It follows the reference's own loop structure -- edge-parallel
// assembly over the edgewise assembly data, ForwardDiff-style full-chunk dual numbers, a CSC matrix whose
// pattern grows through _addnz (exact zeros are not inserted) -- and NOT the row-gather layout of the CUDA
// path, so that agreement between the two is a real check.
//
//   form factors ........ src/vfvm_formfactors.jl:12-332          -> cellfactors(), bfacefactors()
//   edgewise factors .... src/vfvm_system.jl:690-755              -> update_grid_edgewise()
//   iterators/_fill! .... src/vfvm_assemblydata.jl:95-119,215-232 -> loops in assemble_nodes/edges
//   dof loops ........... src/vfvm_assemblydata.jl:245-302,350-382-> inlined in assemble_*()
//   assembly ............ src/vfvm_assembly.jl:9-28,38-126,128-200,318-407,520-643
//   evaluators .......... src/vfvm_physics.jl:335-357,421-452     -> eval_* with vo::Dual<P>
//   Dirichlet init ...... src/vfvm_system.jl:947-1042             -> vo_initialize()
//
// Third-party behaviour that is NOT under /root/reference and is restated from the published algorithm:
//   ExtendableGrids 1.11 (Project.toml:58): edge enumeration = unique node pairs in CSC order of the lower
//     triangle of the node-node adjacency (sorted by smaller node, then larger node), edgenodes[1] = larger
//     node; local_celledgenodes: Edge1D (1,2); Triangle2D (2,3),(3,1),(1,2); Tetrahedron3D (1,2),(1,3),(1,4),
//     (2,3),(2,4),(3,4) (consistent with the index tables at src/vfvm_formfactors.jl:170-176).  Edge order is
//     asserted by no reference test => "parity unpinned" for edge numbering; results do not depend on it.
//   ExtendableSparse 1.6/2 (Project.toml:59): rawupdateindex!(A,+,v,i,j) adds into an existing CSC entry or
//     appends to a per-column extension that flush! merges (sorted rows per column).
//   ForwardDiff: see dual.hpp.
//
// Parity status: PINNED against the reference's own known answers (tests/test_oracle_golden.py):
//   Example301 solution[43], Example207 U[15], Example410 norm, Example105/106/107/110/210/215 values,
//   test010 Bernoulli accuracy, test020 2D==3D-face form factors, test030 1D Laplace, test120 node volumes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <type_traits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "physics.hpp"

namespace vo {

static const double DIRICHLET = 1.0e30;  // src/vfvm_system.jl:329

// ------------------------------------------------------------------------------------------------ grid
struct Grid {
    int dim = 0, coordsys = 0, N = 0, C = 0, NB = 0;
    std::vector<double> coord;
    std::vector<int> cellnodes, cellregions, bfacenodes, bfaceregions;
    int ncellregions = 0, nbfaceregions = 0;
    // derived, ExtendableGrids CellEdges / EdgeNodes
    int E = 0;
    std::vector<int> edgenodes, celledges;
    int nn() const { return dim + 1; }
    int ne() const { return dim * (dim + 1) / 2; }
};

static const int LEN1[1][2] = {{0, 1}};
static const int LEN2[3][2] = {{1, 2}, {2, 0}, {0, 1}};
static const int LEN3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static inline const int (*local_celledgenodes(int dim))[2] {
    return dim == 1 ? LEN1 : (dim == 2 ? LEN2 : LEN3);
}

static void prepare_edges(Grid& g) {
    const int nn = g.nn(), ne = g.ne();
    const int(*len)[2] = local_celledgenodes(g.dim);
    std::vector<uint64_t> keys((size_t)g.C * ne);
    for (int c = 0; c < g.C; c++)
        for (int ie = 0; ie < ne; ie++) {
            int a = g.cellnodes[(size_t)c * nn + len[ie][0]], b = g.cellnodes[(size_t)c * nn + len[ie][1]];
            uint64_t lo = std::min(a, b), hi = std::max(a, b);
            keys[(size_t)c * ne + ie] = (lo << 32) | hi;  // sort by smaller node (CSC column), then larger (row)
        }
    std::vector<uint64_t> uniq = keys;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    g.E = (int)uniq.size();
    g.edgenodes.resize((size_t)2 * g.E);
    for (int e = 0; e < g.E; e++) {
        g.edgenodes[2 * (size_t)e + 0] = (int)(uniq[e] & 0xffffffffu);  // larger node
        g.edgenodes[2 * (size_t)e + 1] = (int)(uniq[e] >> 32);         // smaller node
    }
    g.celledges.resize((size_t)g.C * ne);
    for (size_t i = 0; i < keys.size(); i++)
        g.celledges[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), keys[i]) - uniq.begin());
}

// ------------------------------------------------------------------------------------------------ form factors
// ExtendableGrids.tricircumcenter! (Shewchuk's formula), used by the cylindrical variant
static void tricircumcenter(double* cc, const double* a, const double* b, const double* c) {
    double xba = b[0] - a[0], yba = b[1] - a[1], xca = c[0] - a[0], yca = c[1] - a[1];
    double balength = xba * xba + yba * yba, calength = xca * xca + yca * yca;
    double denominator = 0.5 / (xba * yca - yba * xca);
    cc[0] = (yca * balength - yba * calength) * denominator + a[0];
    cc[1] = (xba * calength - xca * balength) * denominator + a[1];
}

// src/vfvm_formfactors.jl:12-235; n = the cell's node ids, npar/epar = node / edge factors
static void cellfactors(int dim, int coordsys, const double* coord, const int* n, double* npar, double* epar) {
    const double PI = 3.14159265358979323846;
    if (dim == 1) {
        double xK = coord[n[0]], xL = coord[n[1]];
        if (coordsys == VFVM_CARTESIAN) {  // :12-23
            double d = std::fabs(xL - xK);
            npar[0] = d / 2;
            npar[1] = d / 2;
            epar[0] = 1 / d;
        } else {
            double r0 = xK, r1 = xL;
            if (r1 < r0) {
                r0 = xL;
                r1 = xK;
            }
            double rhalf = 0.5 * (r1 + r0);
            if (coordsys == VFVM_CYLINDRICAL) {  // Polar1D :25-43
                npar[0] = PI * (rhalf * rhalf - r0 * r0);
                npar[1] = PI * (r1 * r1 - rhalf * rhalf);
                epar[0] = 2.0 * PI * rhalf / (r1 - r0);
            } else {  // Spherical1D :45-62
                npar[0] = PI * (rhalf * rhalf * rhalf - r0 * r0 * r0) * 4.0 / 3.0;
                npar[1] = PI * (r1 * r1 * r1 - rhalf * rhalf * rhalf) * 4.0 / 3.0;
                epar[0] = 4.0 * PI * (rhalf * rhalf) / (r1 - r0);
            }
        }
        return;
    }
    if (dim == 2) {  // :64-161
        const int(*en)[2] = LEN2;
        double V[2][3];
        for (int i = 0; i < 3; i++)
            for (int d = 0; d < 2; d++) V[d][i] = coord[2 * (size_t)n[en[i][0]] + d] - coord[2 * (size_t)n[en[i][1]] + d];
        double det = V[0][2] * V[1][1] - V[0][1] * V[1][2];
        double vol = std::fabs(0.5 * det);
        double ivol = 1.0 / vol;
        double dd[3];
        for (int i = 0; i < 3; i++) dd[i] = V[0][i] * V[0][i] + V[1][i] * V[1][i];
        double rcc = 0.0, emid[3] = {0, 0, 0};
        if (coordsys == VFVM_CYLINDRICAL) {
            for (int i = 0; i < 3; i++) emid[i] = 0.5 * (coord[2 * (size_t)n[en[i][0]]] + coord[2 * (size_t)n[en[i][1]]]);
            double cc[2];
            tricircumcenter(cc, &coord[2 * (size_t)n[0]], &coord[2 * (size_t)n[1]], &coord[2 * (size_t)n[2]]);
            rcc = cc[0];
        }
        epar[0] = (dd[1] + dd[2] - dd[0]) * 0.125 * ivol;
        epar[1] = (dd[2] + dd[0] - dd[1]) * 0.125 * ivol;
        epar[2] = (dd[0] + dd[1] - dd[2]) * 0.125 * ivol;
        npar[0] = npar[1] = npar[2] = 0.0;
        if (coordsys == VFVM_CARTESIAN) {
            for (int i = 0; i < 3; i++) {
                npar[en[i][0]] += epar[i] * dd[i] * 0.25;
                npar[en[i][1]] += epar[i] * dd[i] * 0.25;
            }
        } else {
            for (int i = 0; i < 3; i++) {
                double r1 = coord[2 * (size_t)n[en[i][0]]], r2 = coord[2 * (size_t)n[en[i][1]]];
                double cylfac1 = 2 * PI * (r1 + rcc + emid[i]) / 3;
                double cylfac2 = 2 * PI * (r2 + rcc + emid[i]) / 3;
                npar[en[i][0]] += epar[i] * dd[i] * 0.25 * cylfac1;
                npar[en[i][1]] += epar[i] * dd[i] * 0.25 * cylfac2;
            }
            for (int i = 0; i < 3; i++) {
                double rmid = (rcc + emid[i]) / 2;
                epar[i] *= 2 * PI * rmid;
            }
        }
        return;
    }
    // dim == 3, :163-235
    static const int pi1[4] = {4, 5, 4, 0}, pi2[4] = {5, 2, 0, 3}, pi3[4] = {3, 1, 2, 1};
    static const int po1[4] = {1, 0, 1, 5}, po2[4] = {0, 3, 5, 2}, po3[4] = {2, 4, 3, 4};
    const int(*en)[2] = LEN3;
    double dd[6];
    for (int i = 0; i < 6; i++) {
        int p1 = n[en[i][0]], p2 = n[en[i][1]];
        double dx = coord[3 * (size_t)p1] - coord[3 * (size_t)p2];
        double dy = coord[3 * (size_t)p1 + 1] - coord[3 * (size_t)p2 + 1];
        double dz = coord[3 * (size_t)p1 + 2] - coord[3 * (size_t)p2 + 2];
        dd[i] = dx * dx + dy * dy + dz * dz;
        epar[i] = 0.0;
    }
    double x1 = coord[3 * (size_t)n[1]] - coord[3 * (size_t)n[0]], y1 = coord[3 * (size_t)n[1] + 1] - coord[3 * (size_t)n[0] + 1],
           z1 = coord[3 * (size_t)n[1] + 2] - coord[3 * (size_t)n[0] + 2];
    double x2 = coord[3 * (size_t)n[2]] - coord[3 * (size_t)n[0]], y2 = coord[3 * (size_t)n[2] + 1] - coord[3 * (size_t)n[0] + 1],
           z2 = coord[3 * (size_t)n[2] + 2] - coord[3 * (size_t)n[0] + 2];
    double x3 = coord[3 * (size_t)n[3]] - coord[3 * (size_t)n[0]], y3 = coord[3 * (size_t)n[3] + 1] - coord[3 * (size_t)n[0] + 1],
           z3 = coord[3 * (size_t)n[3] + 2] - coord[3 * (size_t)n[0] + 2];
    double det = (x1 * (y2 * z3 - y3 * z2) + x2 * (y3 * z1 - y1 * z3) + x3 * (y1 * z2 - y2 * z1));
    if (det < 0) det = -det;
    double vol = det / 6;
    double vv = 96 * 6 * vol;
    for (int i = 0; i < 4; i++) {
        npar[i] = 0.0;
        int i1 = pi1[i], i2 = pi2[i], i3 = pi3[i];
        double h1 = dd[i1] * (dd[i2] + dd[i3] - dd[i1]);
        double h2 = dd[i2] * (dd[i3] + dd[i1] - dd[i2]);
        double h3 = dd[i3] * (dd[i1] + dd[i2] - dd[i3]);
        double df = h1 + h2 + h3;
        double vf = (h1 * dd[po1[i]] + h2 * dd[po2[i]] + h3 * dd[po3[i]] - 2 * dd[i1] * dd[i2] * dd[i3]) / (vv * df);
        epar[i1] += h1 * vf;
        epar[i2] += h2 * vf;
        epar[i3] += h3 * vf;
    }
    for (int i = 0; i < 6; i++) {
        npar[en[i][0]] += epar[i];
        npar[en[i][1]] += epar[i];
        epar[i] = 6 * epar[i] / dd[i];
    }
}

// src/vfvm_formfactors.jl:244-332; n = the boundary face's node ids
static void bfacefactors(int dim, int coordsys, const double* coord, const int* n, double* npar, double* epar) {
    const double PI = 3.14159265358979323846;
    if (dim == 1) {  // Vertex0D :244-261
        double r = coord[n[0]];
        if (coordsys == VFVM_CARTESIAN) npar[0] = 1.0;
        else if (coordsys == VFVM_CYLINDRICAL) npar[0] = 2 * PI * r;
        else npar[0] = 4 * PI * (r * r);
        return;
    }
    if (dim == 2) {  // Edge1D :263-291
        int i1 = n[0], i2 = n[1];
        if (coordsys == VFVM_CARTESIAN) {
            double dx = coord[2 * (size_t)i1] - coord[2 * (size_t)i2], dy = coord[2 * (size_t)i1 + 1] - coord[2 * (size_t)i2 + 1];
            double d = std::sqrt(dx * dx + dy * dy);
            npar[0] = d / 2;
            npar[1] = d / 2;
            epar[0] = 1 / d;
        } else {
            double r1 = coord[2 * (size_t)i1], r2 = coord[2 * (size_t)i2], z1 = coord[2 * (size_t)i1 + 1], z2 = coord[2 * (size_t)i2 + 1];
            double dr = r1 - r2, rmid = (r1 + r2) / 2, dz = z1 - z2;
            double l = std::sqrt(dr * dr + dz * dz);
            npar[0] = PI * (r1 + rmid) * l / 2;
            npar[1] = PI * (r2 + rmid) * l / 2;
            epar[0] = 0.0;
        }
        return;
    }
    // Triangle2D in Cartesian3D :293-332
    const int(*en)[2] = LEN2;
    epar[0] = epar[1] = epar[2] = 0.0;
    for (int j = 0; j < 3; j++) {
        double d = coord[3 * (size_t)n[en[0][0]] + j] - coord[3 * (size_t)n[en[0][1]] + j];
        epar[0] += d * d;
        d = coord[3 * (size_t)n[en[1][0]] + j] - coord[3 * (size_t)n[en[1][1]] + j];
        epar[1] += d * d;
        d = coord[3 * (size_t)n[en[2][0]] + j] - coord[3 * (size_t)n[en[2][1]] + j];
        epar[2] += d * d;
    }
    double dd[3] = {epar[0], epar[1], epar[2]};
    epar[0] = (dd[1] + dd[2] - dd[0]) * dd[0];
    epar[1] = (dd[2] + dd[0] - dd[1]) * dd[1];
    epar[2] = (dd[0] + dd[1] - dd[2]) * dd[2];
    double vol = std::sqrt(epar[0] + epar[1] + epar[2]) * 0.25;
    double d = 1.0 / (8 * vol);
    npar[0] = npar[1] = npar[2] = 0.0;
    for (int i = 0; i < 3; i++) {
        npar[en[i][0]] += epar[i] * d * 0.25;
        npar[en[i][1]] += epar[i] * d * 0.25;
    }
    epar[0] = epar[0] * d / dd[0];
    epar[1] = epar[1] * d / dd[1];
    epar[2] = epar[2] * d / dd[2];
}

// nregions x nitems CSC (one column per node / edge): src/vfvm_assemblydata.jl:39-55
struct FactorCSC {
    std::vector<int64_t> colptr;
    std::vector<int> region;  // rowval, 1-based region label
    std::vector<double> fac;  // nzval
};

// ------------------------------------------------------------------------------------------------ matrix
// ExtendableSparseMatrixCSC restated: CSC + per-column extension merged by flush()
struct ExtMatrix {
    int n = 0;
    std::vector<int64_t> colptr;
    std::vector<int> rowval;
    std::vector<double> nzval;
    std::vector<std::vector<std::pair<int, double>>> ext;
    void init(int n_) {
        n = n_;
        colptr.assign((size_t)n + 1, 0);
        rowval.clear();
        nzval.clear();
        ext.assign((size_t)n, {});
    }
    void zero() {  // zero!(matrix) src/vfvm_assembly.jl:31-34 (after flush the extension is empty)
        std::fill(nzval.begin(), nzval.end(), 0.0);
        for (auto& e : ext)
            for (auto& p : e) p.second = 0.0;
    }
    inline void update(int i, int j, double v) {  // rawupdateindex!(A, +, v, i, j)
        int64_t lo = colptr[j], hi = colptr[j + 1];
        const int* rv = rowval.data();
        while (lo < hi) {  // binary search in column j
            int64_t mid = (lo + hi) >> 1;
            if (rv[mid] < i) lo = mid + 1;
            else hi = mid;
        }
        if (lo < colptr[j + 1] && rv[lo] == i) {
            nzval[lo] += v;
            return;
        }
        for (auto& p : ext[j])
            if (p.first == i) {
                p.second += v;
                return;
            }
        ext[j].emplace_back(i, v);
    }
    int64_t nnznew() const {
        int64_t s = 0;
        for (auto& e : ext) s += (int64_t)e.size();
        return s;
    }
    void flush() {
        if (nnznew() == 0) return;
        std::vector<int64_t> ncolptr((size_t)n + 1, 0);
        for (int j = 0; j < n; j++) ncolptr[j + 1] = ncolptr[j] + (colptr[j + 1] - colptr[j]) + (int64_t)ext[j].size();
        std::vector<int> nrow((size_t)ncolptr[n]);
        std::vector<double> nval((size_t)ncolptr[n]);
        for (int j = 0; j < n; j++) {
            auto& e = ext[j];
            std::sort(e.begin(), e.end());
            int64_t a = colptr[j], ae = colptr[j + 1], o = ncolptr[j];
            size_t b = 0;
            while (a < ae || b < e.size()) {
                if (b >= e.size() || (a < ae && rowval[a] < e[b].first)) {
                    nrow[o] = rowval[a];
                    nval[o] = nzval[a];
                    a++;
                } else {
                    nrow[o] = e[b].first;
                    nval[o] = e[b].second;
                    b++;
                }
                o++;
            }
            e.clear();
            e.shrink_to_fit();
        }
        colptr.swap(ncolptr);
        rowval.swap(nrow);
        nzval.swap(nval);
    }
};

// ------------------------------------------------------------------------------------------------ system
struct System {
    Grid g;
    FactorCSC nodefactors, edgefactors;
    std::vector<double> bfacenodefactors, bfaceedgefactors;
    int n = 0;                                // nspecies
    std::vector<uint8_t> region_species;      // n x ncellregions
    std::vector<uint8_t> bregion_species;     // n x nbregions_bs (enable_boundary_species!, src/vfvm_system.jl:492-515), may be empty
    int nbregions_bs = 0;
    std::vector<uint8_t> node_dof;            // n x N
    bool species_homogeneous = true;
    Physics ph;
    std::vector<double> boundary_factors, boundary_values;  // n x nbfaceregions
    ExtMatrix A;
    bool nan_seen = false;
    // node-range partitions for the threaded loop (src/vfvm_assembly.jl:571-593)
    std::vector<int> part_nodes, part_edges;  // partition boundaries
    std::string err;
};

static void build_factor_csc(int nreg, int nitems, const std::vector<double>& acc, const std::vector<uint8_t>& touched, FactorCSC& out) {
    out.colptr.assign((size_t)nitems + 1, 0);
    for (int i = 0; i < nitems; i++) {
        int c = 0;
        for (int r = 0; r < nreg; r++) c += touched[(size_t)i * nreg + r];
        out.colptr[i + 1] = out.colptr[i] + c;
    }
    out.region.resize((size_t)out.colptr[nitems]);
    out.fac.resize((size_t)out.colptr[nitems]);
    for (int i = 0; i < nitems; i++) {
        int64_t o = out.colptr[i];
        for (int r = 0; r < nreg; r++)
            if (touched[(size_t)i * nreg + r]) {
                out.region[o] = r + 1;
                out.fac[o] = acc[(size_t)i * nreg + r];
                o++;
            }
    }
}

// src/vfvm_system.jl:690-755
static void update_grid_edgewise(System& s) {
    Grid& g = s.g;
    prepare_edges(g);
    const int nn = g.nn(), ne = g.ne(), nreg = g.ncellregions;
    std::vector<double> cnf((size_t)g.N * nreg, 0.0), cef((size_t)g.E * nreg, 0.0);
    std::vector<uint8_t> tn((size_t)g.N * nreg, 0), te((size_t)g.E * nreg, 0);
    double npar[4], epar[6];
    for (int c = 0; c < g.C; c++) {  // :716-726, accumulation in cell order
        cellfactors(g.dim, g.coordsys, g.coord.data(), &g.cellnodes[(size_t)c * nn], npar, epar);
        int ireg = g.cellregions[c] - 1;
        for (int i = 0; i < nn; i++) {
            size_t k = (size_t)g.cellnodes[(size_t)c * nn + i] * nreg + ireg;
            cnf[k] += npar[i];
            tn[k] = 1;
        }
        for (int i = 0; i < ne; i++) {
            size_t k = (size_t)g.celledges[(size_t)c * ne + i] * nreg + ireg;
            cef[k] += epar[i];
            te[k] = 1;
        }
    }
    build_factor_csc(nreg, g.N, cnf, tn, s.nodefactors);
    build_factor_csc(nreg, g.E, cef, te, s.edgefactors);
    const int nbn = g.dim, nbe = (g.dim == 3) ? 3 : (g.dim == 2 ? 1 : 0);
    s.bfacenodefactors.assign((size_t)g.NB * nbn, 0.0);
    s.bfaceedgefactors.assign((size_t)g.NB * std::max(nbe, 1), 0.0);
    double bn[3], be[3];
    for (int b = 0; b < g.NB; b++) {  // :730-735
        bfacefactors(g.dim, g.coordsys, g.coord.data(), &g.bfacenodes[(size_t)b * nbn], bn, be);
        for (int i = 0; i < nbn; i++) s.bfacenodefactors[(size_t)b * nbn + i] = bn[i];
        for (int i = 0; i < nbe; i++) s.bfaceedgefactors[(size_t)b * nbe + i] = be[i];
    }
}

// enable_species! src/vfvm_system.jl:433-459 + species_homogeneous scan :539-549
static void complete_species(System& s) {
    Grid& g = s.g;
    const int n = s.n, nn = g.nn();
    s.node_dof.assign((size_t)g.N * n, 0);
    for (int c = 0; c < g.C; c++) {
        int ireg = g.cellregions[c] - 1;
        for (int i = 0; i < n; i++)
            if (s.region_species[(size_t)ireg * n + i])
                for (int l = 0; l < nn; l++) s.node_dof[(size_t)g.cellnodes[(size_t)c * nn + l] * n + i] = 1;
    }
    if (!s.bregion_species.empty()) {  // src/vfvm_system.jl:502-513
        const int nbn = g.dim;
        for (int b = 0; b < g.NB; b++) {
            const int br = g.bfaceregions[b] - 1;
            if (br >= s.nbregions_bs) continue;
            for (int i = 0; i < n; i++)
                if (s.bregion_species[(size_t)br * n + i])
                    for (int l = 0; l < nbn; l++) s.node_dof[(size_t)g.bfacenodes[(size_t)b * nbn + l] * n + i] = 1;
        }
    }
    s.species_homogeneous = true;
    for (auto v : s.node_dof)
        if (!v) s.species_homogeneous = false;
}

// ------------------------------------------------------------------------------------------------ assembly
static inline void addnz(System& s, int i, int j, double v, double fac) {  // src/vfvm_assembly.jl:21-28
    if (std::isnan(v)) {
        s.nan_seen = true;
        return;
    }
    if (v != 0.0) s.A.update(i, j, v * fac);
}

template <int NS>
static void assemble_nodes(System& s, int n0, int n1, const double* U, const double* UOld, double* F, double time, double tstepinv, double lambda) {
    const Grid& g = s.g;
    const int n = NS;
    NodeCtx node;
    node.dim = g.dim;
    node.time = time;
    node.embed = lambda;
    typedef Dual<NS> D;
    double src[NS], oldstor[NS], uo[NS];
    D UK[NS], rea[NS], stor[NS];
    for (int K = n0; K < n1; K++) {
        for (int64_t k = s.nodefactors.colptr[K]; k < s.nodefactors.colptr[K + 1]; k++) {
            node.index = K;  // _fill! src/vfvm_assemblydata.jl:215-219
            node.region = s.nodefactors.region[k];
            node.fac = s.nodefactors.fac[k];
            node.x = &g.coord[(size_t)K * g.dim];
            for (int i = 0; i < n; i++) {
                UK[i] = D(U[(size_t)K * n + i]);
                UK[i].d[i] = 1.0;
                uo[i] = UOld[(size_t)K * n + i];
                src[i] = 0.0;
                oldstor[i] = 0.0;
                rea[i] = D(0.0);
                stor[i] = D(0.0);
            }
            eval_source(s.ph.slot[VFVM_SLOT_SOURCE], n, src, node, s.ph.nodal_source);
            eval_reaction(s.ph.slot[VFVM_SLOT_REACTION], n, rea, UK, node);
            eval_storage(s.ph.slot[VFVM_SLOT_STORAGE], n, stor, UK, node);
            eval_storage(s.ph.slot[VFVM_SLOT_STORAGE], n, oldstor, uo, node);
            const uint8_t* rs = &s.region_species[(size_t)(node.region - 1) * n];
            for (int i = 0; i < n; i++) {  // assemble_res_jac(node) src/vfvm_assemblydata.jl:245-270
                if (!rs[i]) continue;
                int idof = K * n + i;
                F[idof] += node.fac * (rea[i].v - src[i] + (stor[i].v - oldstor[i]) * tstepinv);  // :95-104
                for (int j = 0; j < n; j++) {
                    if (!rs[j]) continue;
                    addnz(s, idof, K * n + j, rea[i].d[j] + stor[i].d[j] * tstepinv, node.fac);  // :106-114
                }
            }
        }
    }
}

template <int NS>
static void assemble_edges(System& s, int e0, int e1, const double* U, double* F, double time, double lambda) {
    const Grid& g = s.g;
    const int n = NS;
    EdgeCtx edge;
    edge.dim = g.dim;
    edge.time = time;
    edge.embed = lambda;
    typedef Dual<2 * NS> D;
    D uK[NS], uL[NS], f[NS];
    for (int ie = e0; ie < e1; ie++) {
        for (int64_t k = s.edgefactors.colptr[ie]; k < s.edgefactors.colptr[ie + 1]; k++) {
            int K = g.edgenodes[2 * (size_t)ie], L = g.edgenodes[2 * (size_t)ie + 1];  // _fill! :226-232
            edge.index = ie;
            edge.nodeK = K;
            edge.nodeL = L;
            edge.region = s.edgefactors.region[k];
            edge.fac = s.edgefactors.fac[k];
            edge.xK = &g.coord[(size_t)K * g.dim];
            edge.xL = &g.coord[(size_t)L * g.dim];
            for (int i = 0; i < n; i++) {  // UKL gather, src/vfvm_assembly.jl:162-163
                uK[i] = D(U[(size_t)K * n + i]);
                uK[i].d[i] = 1.0;
                uL[i] = D(U[(size_t)L * n + i]);
                uL[i].d[n + i] = 1.0;
                f[i] = D(0.0);
            }
            eval_flux(s.ph.slot[VFVM_SLOT_FLUX], n, f, uK, uL, edge);
            const uint8_t* rs = &s.region_species[(size_t)(edge.region - 1) * n];
            for (int i = 0; i < n; i++) {  // assemble_res_jac(edge) src/vfvm_assemblydata.jl:350-382
                if (!rs[i]) continue;
                int idofK = K * n + i, idofL = L * n + i;
                double val = edge.fac * f[i].v;  // src/vfvm_assembly.jl:169-173
                F[idofK] += val;
                F[idofL] += -val;
                for (int j = 0; j < n; j++) {
                    if (!rs[j]) continue;
                    int jdofK = K * n + j, jdofL = L * n + j;  // :175-192
                    addnz(s, idofK, jdofK, +f[i].d[j], edge.fac);
                    addnz(s, idofL, jdofK, -f[i].d[j], edge.fac);
                    addnz(s, idofK, jdofL, +f[i].d[j + n], edge.fac);
                    addnz(s, idofL, jdofL, -f[i].d[j + n], edge.fac);
                }
            }
            if (s.ph.slot[VFVM_SLOT_EDGEREACTION].id != VFVM_NONE) {  // src/vfvm_assembly.jl:202-239
                for (int i = 0; i < n; i++) f[i] = D(0.0);
                eval_edgereaction(s.ph.slot[VFVM_SLOT_EDGEREACTION], n, f, uK, uL, edge);
                for (int i = 0; i < n; i++) {
                    if (!rs[i]) continue;
                    int idofK = K * n + i, idofL = L * n + i;
                    double val = edge.fac * f[i].v;  // :207-211: the same sign at both ends
                    F[idofK] += val;
                    F[idofL] += val;
                    for (int j = 0; j < n; j++) {
                        if (!rs[j]) continue;
                        int jdofK = K * n + j, jdofL = L * n + j;  // :213-230, signs as the reference has them
                        addnz(s, idofK, jdofK, +f[i].d[j], edge.fac);
                        addnz(s, idofL, jdofK, -f[i].d[j], edge.fac);
                        addnz(s, idofK, jdofL, -f[i].d[j + n], edge.fac);
                        addnz(s, idofL, jdofL, +f[i].d[j + n], edge.fac);
                    }
                }
            }
        }
    }
}

template <int NS>
static void assemble_bnodes(System& s, const double* U, const double* UOld, double* F, double time, double tstepinv, double lambda) {
    const Grid& g = s.g;
    const int n = NS, nbn = g.dim;
    bool has_legacy_bc = false;  // src/vfvm_assembly.jl:329
    for (double v : s.boundary_factors) has_legacy_bc |= (v != 0.0);
    for (double v : s.boundary_values) has_legacy_bc |= (v != 0.0);
    BNodeCtx b;
    b.dim = g.dim;
    b.time = time;
    b.embed = lambda;
    typedef Dual<NS> D;
    D UK[NS], res[NS];
    double dv[NS];
    b.dirichlet_value = dv;
    for (int ibf = 0; ibf < g.NB; ibf++) {
        for (int ibn = 0; ibn < nbn; ibn++) {
            b.ibface = ibf;  // _fill! src/vfvm_assemblydata.jl:143-160
            b.ibnode = ibn;
            b.region = g.bfaceregions[ibf];
            b.index = g.bfacenodes[(size_t)ibf * nbn + ibn];
            b.fac = s.bfacenodefactors[(size_t)ibf * nbn + ibn];
            b.x = &g.coord[(size_t)b.index * g.dim];
            b.Dirichlet = DIRICHLET;
            int K = b.index;
            if (has_legacy_bc) {  // src/vfvm_assembly.jl:355-387
                b.Dirichlet = DIRICHLET / b.fac;
                for (int i = 0; i < n; i++) {
                    int idof = K * n + i;
                    double bf = s.boundary_factors[(size_t)(b.region - 1) * n + i];
                    double bv = s.boundary_values[(size_t)(b.region - 1) * n + i];
                    if (bf == DIRICHLET) {
                        F[idof] += bf * (U[idof] - bv);
                        addnz(s, idof, idof, bf, 1.0);
                    } else {
                        F[idof] += b.fac * (bf * U[idof] - bv);
                        addnz(s, idof, idof, bf, b.fac);
                    }
                }
            }
            for (int i = 0; i < n; i++) {
                UK[i] = D(U[(size_t)K * n + i]);
                UK[i].d[i] = 1.0;
                res[i] = D(0.0);
            }
            eval_breaction(s.ph, n, res, UK, b);  // bsource is not registered (always zero)
            for (int i = 0; i < n; i++) {  // assemble_res_jac(bnode) src/vfvm_assemblydata.jl:278-302
                if (!s.node_dof[(size_t)K * n + i]) continue;
                int idof = K * n + i;
                F[idof] += b.fac * res[i].v;  // :399
                for (int j = 0; j < n; j++) {
                    if (!s.node_dof[(size_t)K * n + j]) continue;
                    addnz(s, idof, K * n + j, res[i].d[j], b.fac);  // :401
                }
            }
            if (s.ph.slot[VFVM_SLOT_BSTORAGE].id != VFVM_NONE) {  // src/vfvm_assembly.jl:409-439
                D st[NS];
                double uo[NS], sto[NS];
                for (int i = 0; i < n; i++) {
                    st[i] = D(0.0);
                    sto[i] = 0.0;
                    uo[i] = UOld[(size_t)K * n + i];
                }
                eval_bstorage(s.ph.slot[VFVM_SLOT_BSTORAGE], n, st, UK, b);
                eval_bstorage(s.ph.slot[VFVM_SLOT_BSTORAGE], n, sto, uo, b);
                for (int i = 0; i < n; i++) {
                    if (!s.node_dof[(size_t)K * n + i]) continue;
                    int idof = K * n + i;
                    F[idof] += b.fac * (st[i].v - sto[i]) * tstepinv;
                    for (int j = 0; j < n; j++) {
                        if (!s.node_dof[(size_t)K * n + j]) continue;
                        addnz(s, idof, K * n + j, st[i].d[j], b.fac * tstepinv);
                    }
                }
            }
        }
    }
}

// node-range partitions: an edge belongs to the partition of its smaller node (edges are sorted by it), so
// partitions p and p+2 share no node as long as every chunk is wider than the graph bandwidth
static void make_partitions(System& s, int nthreads) {
    const Grid& g = s.g;
    int bw = 0;
    for (int e = 0; e < g.E; e++) bw = std::max(bw, g.edgenodes[2 * (size_t)e] - g.edgenodes[2 * (size_t)e + 1]);
    int maxparts = std::max(1, g.N / std::max(1, bw + 1));
    int nparts = std::min(maxparts, std::max(1, 4 * nthreads));
    if (nparts < 2) nparts = 1;
    s.part_nodes.assign((size_t)nparts + 1, 0);
    for (int p = 0; p <= nparts; p++) s.part_nodes[p] = (int)((int64_t)g.N * p / nparts);
    s.part_edges.assign((size_t)nparts + 1, 0);
    int e = 0;
    for (int p = 0; p < nparts; p++) {
        s.part_edges[p] = e;
        while (e < g.E && g.edgenodes[2 * (size_t)e + 1] < s.part_nodes[p + 1]) e++;
    }
    s.part_edges[nparts] = g.E;
}

// eval_and_assemble src/vfvm_assembly.jl:520-643
template <int NS>
static int eval_and_assemble(System& s, const double* U, const double* UOld, double* F, double time, double tstep, double lambda, int nthreads) {
    const Grid& g = s.g;
    const int n = NS;
    s.A.zero();  // :542-543
    std::fill(F, F + (size_t)g.N * n, 0.0);
    double tstepinv = 1.0 / tstep;  // :554
    s.nan_seen = false;
    if (nthreads <= 1) {  // :561-570
        assemble_nodes<NS>(s, 0, g.N, U, UOld, F, time, tstepinv, lambda);
        assemble_edges<NS>(s, 0, g.E, U, F, time, lambda);
    } else {  // :571-583: colours sequential, partitions of one colour in parallel
        if (s.part_nodes.empty()) make_partitions(s, nthreads);
        int nparts = (int)s.part_nodes.size() - 1;
        for (int color = 0; color < 2; color++) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
            for (int p = color; p < nparts; p += 2)
                assemble_nodes<NS>(s, s.part_nodes[p], s.part_nodes[p + 1], U, UOld, F, time, tstepinv, lambda);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
            for (int p = color; p < nparts; p += 2) assemble_edges<NS>(s, s.part_edges[p], s.part_edges[p + 1], U, F, time, lambda);
        }
    }
    assemble_bnodes<NS>(s, U, UOld, F, time, tstepinv, lambda);  // boundary loop is a single partition, src/vfvm_system.jl:751-754
    if (!s.species_homogeneous) {                // src/vfvm_system.jl:1012-1026
        for (int K = 0; K < g.N; K++)
            for (int i = 0; i < n; i++)
                if (!s.node_dof[(size_t)K * n + i]) {
                    int idof = K * n + i;
                    F[idof] += U[idof] - UOld[idof];
                    s.A.update(idof, idof, 1.0);
                }
    }
    s.A.flush();
    return s.nan_seen ? VFVM_ERR_NAN : 0;
}

// _initialize_dirichlet! + _initialize_inactive_dof! src/vfvm_system.jl:947-1042
template <int NS>
static void initialize(System& s, double* U, double time, double lambda) {
    const Grid& g = s.g;
    const int n = NS, nbn = g.dim;
    BNodeCtx b;
    b.dim = g.dim;
    b.time = time;
    b.embed = lambda;
    double dv[NS], y[NS], u[NS];
    b.dirichlet_value = dv;
    for (int ibf = 0; ibf < g.NB; ibf++)
        for (int ibn = 0; ibn < nbn; ibn++) {
            b.ibface = ibf;
            b.ibnode = ibn;
            b.region = g.bfaceregions[ibf];
            b.index = g.bfacenodes[(size_t)ibf * nbn + ibn];
            b.fac = s.bfacenodefactors[(size_t)ibf * nbn + ibn];
            b.x = &g.coord[(size_t)b.index * g.dim];
            b.Dirichlet = DIRICHLET;
            for (int i = 0; i < n; i++) {
                dv[i] = INFINITY;
                y[i] = 0.0;
                u[i] = U[(size_t)b.index * n + i];
            }
            eval_breaction(s.ph, n, y, u, b);
            for (int i = 0; i < n; i++) {
                if (!std::isinf(dv[i])) U[(size_t)b.index * n + i] = dv[i];
                if (!s.boundary_factors.empty()) {
                    double bf = s.boundary_factors[(size_t)(b.region - 1) * n + i];
                    // isapprox(bf, 1e30): |a-b| <= sqrt(eps) * max(|a|,|b|)
                    if (std::fabs(bf - DIRICHLET) <= 1.4901161193847656e-8 * std::max(std::fabs(bf), DIRICHLET))
                        U[(size_t)b.index * n + i] = s.boundary_values[(size_t)(b.region - 1) * n + i];
                }
            }
        }
    if (!s.species_homogeneous)
        for (size_t k = 0; k < s.node_dof.size(); k++)
            if (!s.node_dof[k]) U[k] = 0.0;
}

#define VO_DISPATCH(NSV, CALL)                    \
    switch (NSV) {                                \
        case 1: { constexpr int NS = 1; CALL; } break;   \
        case 2: { constexpr int NS = 2; CALL; } break;   \
        case 3: { constexpr int NS = 3; CALL; } break;   \
        case 4: { constexpr int NS = 4; CALL; } break;   \
        case 5: { constexpr int NS = 5; CALL; } break;   \
        case 6: { constexpr int NS = 6; CALL; } break;   \
        case 8: { constexpr int NS = 8; CALL; } break;   \
        case 10: { constexpr int NS = 10; CALL; } break; \
        case 50: { constexpr int NS = 50; CALL; } break; \
        default: return VFVM_ERR_UNSUPPORTED;     \
    }

}  // namespace vo

// ================================================================================================ C interface (ctypes)
using namespace vo;

extern "C" {

void* vo_create() { return new System(); }
void vo_destroy(void* h) { delete (System*)h; }

int vo_set_grid(void* h, int dim, int coordsys, int N, int C, int NB, const double* coord, const int* cellnodes, const int* cellregions,
                const int* bfacenodes, const int* bfaceregions) {
    System& s = *(System*)h;
    Grid& g = s.g;
    if (dim < 1 || dim > 3) return VFVM_ERR_ARG;
    g.dim = dim;
    g.coordsys = coordsys;
    g.N = N;
    g.C = C;
    g.NB = NB;
    g.coord.assign(coord, coord + (size_t)dim * N);
    g.cellnodes.assign(cellnodes, cellnodes + (size_t)(dim + 1) * C);
    g.cellregions.assign(cellregions, cellregions + C);
    g.bfacenodes.assign(bfacenodes, bfacenodes + (size_t)dim * NB);
    g.bfaceregions.assign(bfaceregions, bfaceregions + NB);
    g.ncellregions = 0;
    for (int r : g.cellregions) g.ncellregions = std::max(g.ncellregions, r);
    g.nbfaceregions = 0;
    for (int r : g.bfaceregions) g.nbfaceregions = std::max(g.nbfaceregions, r);
    return 0;
}

int vo_update_grid(void* h) {
    update_grid_edgewise(*(System*)h);
    return 0;
}
int vo_num_edges(void* h) { return ((System*)h)->g.E; }
void vo_get_edgenodes(void* h, int* out) {
    auto& v = ((System*)h)->g.edgenodes;
    std::copy(v.begin(), v.end(), out);
}
void vo_get_celledges(void* h, int* out) {
    auto& v = ((System*)h)->g.celledges;
    std::copy(v.begin(), v.end(), out);
}
int64_t vo_num_nodefactors(void* h) { return (int64_t)((System*)h)->nodefactors.fac.size(); }
int64_t vo_num_edgefactors(void* h) { return (int64_t)((System*)h)->edgefactors.fac.size(); }
static void copy_factors(const FactorCSC& f, int64_t* colptr, int* region, double* fac) {
    std::copy(f.colptr.begin(), f.colptr.end(), colptr);
    std::copy(f.region.begin(), f.region.end(), region);
    std::copy(f.fac.begin(), f.fac.end(), fac);
}
void vo_get_nodefactors(void* h, int64_t* colptr, int* region, double* fac) { copy_factors(((System*)h)->nodefactors, colptr, region, fac); }
void vo_get_edgefactors(void* h, int64_t* colptr, int* region, double* fac) { copy_factors(((System*)h)->edgefactors, colptr, region, fac); }
void vo_get_bfacefactors(void* h, double* out) {
    auto& v = ((System*)h)->bfacenodefactors;
    std::copy(v.begin(), v.end(), out);
}

int vo_set_system(void* h, int nspecies, const uint8_t* region_species) {
    System& s = *(System*)h;
    s.n = nspecies;
    size_t sz = (size_t)nspecies * s.g.ncellregions;
    if (region_species) s.region_species.assign(region_species, region_species + sz);
    else s.region_species.assign(sz, 1);
    complete_species(s);
    s.boundary_factors.assign((size_t)nspecies * s.g.nbfaceregions, 0.0);
    s.boundary_values.assign((size_t)nspecies * s.g.nbfaceregions, 0.0);
    s.A.init(nspecies * s.g.N);
    s.part_nodes.clear();
    return 0;
}
int vo_set_boundary_species(void* h, int nbregions, const uint8_t* bregion_species) {
    System& s = *(System*)h;
    s.nbregions_bs = bregion_species ? nbregions : 0;
    if (bregion_species) s.bregion_species.assign(bregion_species, bregion_species + (size_t)s.n * nbregions);
    else s.bregion_species.clear();
    complete_species(s);
    return 0;
}
int vo_set_physics(void* h, int slot, int id, const double* params, int np) {
    System& s = *(System*)h;
    if (slot < 0 || slot >= VFVM_NUM_SLOTS) return VFVM_ERR_ARG;
    s.ph.slot[slot].id = id;
    s.ph.slot[slot].p.assign(params, params + np);
    return 0;
}
int vo_set_nodal_source(void* h, const double* table) {
    System& s = *(System*)h;
    s.ph.nodal_source.assign(table, table + (size_t)s.n * s.g.N);
    return 0;
}
int vo_set_legacy_bc(void* h, int nbregions, const double* factors, const double* values) {
    System& s = *(System*)h;
    if (nbregions < s.g.nbfaceregions) return VFVM_ERR_ARG;  // a rank-local grid piece may lack faces of the highest regions
    s.boundary_factors.assign(factors, factors + (size_t)s.n * nbregions);
    s.boundary_values.assign(values, values + (size_t)s.n * nbregions);
    return 0;
}
int vo_set_bc_entries(void* h, int nentries, const vfvm_bc_entry* e) {
    System& s = *(System*)h;
    s.ph.bc.clear();
    for (int i = 0; i < nentries; i++) s.ph.bc.push_back({e[i].kind, e[i].species, e[i].region, e[i].has_ramp, e[i].value, e[i].factor, e[i].t0, e[i].t1, e[i].v0, e[i].v1});
    return 0;
}

int vo_assemble(void* h, const double* U, const double* UOld, double* F, double time, double tstep, double lambda, int nthreads) {
    System& s = *(System*)h;
    int rc = 0;
    VO_DISPATCH(s.n, rc = eval_and_assemble<NS>(s, U, UOld, F, time, tstep, lambda, nthreads));
    return rc;
}
int vo_initialize(void* h, double* U, double time, double lambda) {
    System& s = *(System*)h;
    VO_DISPATCH(s.n, initialize<NS>(s, U, time, lambda));
    return 0;
}
int64_t vo_matrix_nnz(void* h) { return (int64_t)((System*)h)->A.nzval.size(); }
void vo_get_matrix_csc(void* h, int64_t* colptr, int64_t* rowval, double* nzval) {
    ExtMatrix& A = ((System*)h)->A;
    std::copy(A.colptr.begin(), A.colptr.end(), colptr);
    std::copy(A.rowval.begin(), A.rowval.end(), rowval);
    std::copy(A.nzval.begin(), A.nzval.end(), nzval);
}
void vo_get_matrix_values(void* h, double* nzval) {
    ExtMatrix& A = ((System*)h)->A;
    std::copy(A.nzval.begin(), A.nzval.end(), nzval);
}

// unit-test probes (test/test010_bernoulli.jl, test/test020_formfactors.jl)
// ---- post-processing integrals, src/vfvm_postprocess.jl:18-67 (integrate) and :109-146 (edgeintegrate) ----------------------
// slot: VFVM_SLOT_REACTION / VFVM_SLOT_STORAGE select the evaluator of a registered node function; id 0 = the identity
// (integrate(system, U), :94-98).  out: n x ncellregions, column-major.
int vo_integrate(void* h, int slot, int id, const double* params, int np, const double* U, double* out) {
    System& s = *(System*)h;
    const Grid& g = s.g;
    const int n = s.n;
    PhysSlot ps;
    ps.id = id;
    ps.p.assign(params, params + np);
    std::fill(out, out + (size_t)n * g.ncellregions, 0.0);
    NodeCtx node;
    node.dim = g.dim;
    std::vector<double> res(n), u(n);
    for (int K = 0; K < g.N; K++)
        for (int64_t k = s.nodefactors.colptr[K]; k < s.nodefactors.colptr[K + 1]; k++) {
            node.index = K;
            node.region = s.nodefactors.region[k];
            node.fac = s.nodefactors.fac[k];
            node.x = &g.coord[(size_t)K * g.dim];
            for (int i = 0; i < n; i++) {
                u[i] = U[(size_t)K * n + i];
                res[i] = 0.0;
            }
            if (id == 0) res = u;
            else if (slot == VFVM_SLOT_STORAGE) eval_storage(ps, n, res.data(), u.data(), node);
            else eval_reaction(ps, n, res.data(), u.data(), node);
            const uint8_t* rs = &s.region_species[(size_t)(node.region - 1) * n];
            for (int i = 0; i < n; i++)
                if (rs[i]) out[(size_t)(node.region - 1) * n + i] += node.fac * res[i];  // :60
        }
    return 0;
}
// integrate(system, F, U; boundary = true), src/vfvm_postprocess.jl:29-46.  out: n x nbfaceregions, column-major.
int vo_integrate_boundary(void* h, int slot, int id, const double* params, int np, const double* U, double* out) {
    System& s = *(System*)h;
    const Grid& g = s.g;
    const int n = s.n, nbn = g.dim;
    Physics tmp;  // the function alone: no boundary-condition helper calls
    if (slot >= 0 && slot < VFVM_NUM_SLOTS) {
        tmp.slot[slot].id = id;
        tmp.slot[slot].p.assign(params, params + np);
    }
    std::fill(out, out + (size_t)n * g.nbfaceregions, 0.0);
    std::vector<double> res(n), u(n);
    BNodeCtx b;
    b.dim = g.dim;
    NodeCtx node;
    node.dim = g.dim;
    for (int ibf = 0; ibf < g.NB; ibf++)
        for (int ibn = 0; ibn < nbn; ibn++) {
            b.ibface = ibf;
            b.ibnode = ibn;
            b.region = g.bfaceregions[ibf];
            b.index = g.bfacenodes[(size_t)ibf * nbn + ibn];
            b.fac = s.bfacenodefactors[(size_t)ibf * nbn + ibn];
            b.x = &g.coord[(size_t)b.index * g.dim];
            b.dirichlet_value = nullptr;
            node.index = b.index;
            node.region = b.region;
            node.x = b.x;
            const int K = b.index;
            for (int i = 0; i < n; i++) {
                u[i] = U[(size_t)K * n + i];
                res[i] = 0.0;
            }
            if (id == 0) res = u;
            else if (slot == VFVM_SLOT_BREACTION) eval_breaction(tmp, n, res.data(), u.data(), b);
            else if (slot == VFVM_SLOT_BSTORAGE) eval_bstorage(tmp.slot[slot], n, res.data(), u.data(), b);
            else if (slot == VFVM_SLOT_STORAGE) eval_storage(tmp.slot[slot], n, res.data(), u.data(), node);
            else eval_reaction(tmp.slot[slot], n, res.data(), u.data(), node);
            for (int i = 0; i < n; i++)
                if (s.node_dof[(size_t)K * n + i]) out[(size_t)(b.region - 1) * n + i] += b.fac * res[i];  // :42, assemble_res(bnode)
        }
    return 0;
}
// id: a registered flux id, or -1 = the W^{1,p} seminorm integrand dim ((u_K - u_L) / h)^p with p = params[0] (:300-312)
int vo_edgeintegrate(void* h, int id, const double* params, int np, const double* U, double* out) {
    System& s = *(System*)h;
    const Grid& g = s.g;
    const int n = s.n;
    PhysSlot ps;
    ps.id = id;
    ps.p.assign(params, params + np);
    std::fill(out, out + (size_t)n * g.ncellregions, 0.0);
    EdgeCtx edge;
    edge.dim = g.dim;
    std::vector<double> res(n), uK(n), uL(n);
    for (int ie = 0; ie < g.E; ie++)
        for (int64_t k = s.edgefactors.colptr[ie]; k < s.edgefactors.colptr[ie + 1]; k++) {
            const int K = g.edgenodes[2 * (size_t)ie], L = g.edgenodes[2 * (size_t)ie + 1];
            edge.index = ie;
            edge.nodeK = K;
            edge.nodeL = L;
            edge.region = s.edgefactors.region[k];
            edge.fac = s.edgefactors.fac[k];
            edge.xK = &g.coord[(size_t)K * g.dim];
            edge.xL = &g.coord[(size_t)L * g.dim];
            double h2 = 0.0;  // meas(edge)^2, src/vfvm_geometryitems.jl (distance of the two edge nodes)
            for (int d = 0; d < g.dim; d++) h2 += (edge.xK[d] - edge.xL[d]) * (edge.xK[d] - edge.xL[d]);
            const double hh = std::sqrt(h2);
            for (int i = 0; i < n; i++) {
                uK[i] = U[(size_t)K * n + i];
                uL[i] = U[(size_t)L * n + i];
                res[i] = 0.0;
            }
            if (id == -1) {
                for (int i = 0; i < n; i++) res[i] = g.dim * std::pow((uK[i] - uL[i]) / hh, params[0]);
            } else if (id == -2) {  // edge average, test/test120_norms.jl:35-38
                for (int i = 0; i < n; i++) res[i] = 0.5 * (uK[i] + uL[i]);
            } else {
                eval_flux(ps, n, res.data(), uK.data(), uL.data(), edge);
            }
            const uint8_t* rs = &s.region_species[(size_t)(edge.region - 1) * n];
            for (int i = 0; i < n; i++)
                if (rs[i]) out[(size_t)(edge.region - 1) * n + i] += hh * hh * edge.fac * res[i] / g.dim;  // :138
        }
    return 0;
}

// flux callback of every edge without form factor: out[e*n + i] = flux(u_K, u_L)_i, K = edge.node[1], L = edge.node[2]
// (the edge loop of nodeflux, src/vfvm_postprocess.jl:191-207)
int vo_edgeflux(void* h, int id, const double* params, int np, const double* U, double* out) {
    System& s = *(System*)h;
    const Grid& g = s.g;
    const int n = s.n;
    PhysSlot ps;
    ps.id = id;
    ps.p.assign(params, params + np);
    EdgeCtx edge;
    edge.dim = g.dim;
    std::vector<double> res(n), uK(n), uL(n);
    for (int ie = 0; ie < g.E; ie++) {
        const int K = g.edgenodes[2 * (size_t)ie], L = g.edgenodes[2 * (size_t)ie + 1];
        edge.index = ie;
        edge.nodeK = K;
        edge.nodeL = L;
        edge.region = 1;
        edge.xK = &g.coord[(size_t)K * g.dim];
        edge.xL = &g.coord[(size_t)L * g.dim];
        for (int i = 0; i < n; i++) {
            uK[i] = U[(size_t)K * n + i];
            uL[i] = U[(size_t)L * n + i];
            res[i] = 0.0;
        }
        eval_flux(ps, n, res.data(), uK.data(), uL.data(), edge);
        for (int i = 0; i < n; i++) out[(size_t)ie * n + i] = res[i];
    }
    return 0;
}

// mass_matrix(state), src/vfvm_diffeq_interface.jl:60-101: Jacobian of the storage term at U = 0, weighted by the node factors.
// out: n x n x N (out[(K*n + i)*n + j] = M[(K,i),(K,j)]); bstorage is not part of the registered library.
int vo_mass_matrix(void* h, double* out) {
    System& s = *(System*)h;
    const Grid& g = s.g;
    const int n = s.n;
    std::fill(out, out + (size_t)n * n * g.N, 0.0);
    NodeCtx node;
    node.dim = g.dim;
    auto run = [&](auto tag) {
        constexpr int NS = decltype(tag)::value;
        typedef Dual<NS> D;
        D u[NS], st[NS];
        for (int K = 0; K < g.N; K++)
            for (int64_t k = s.nodefactors.colptr[K]; k < s.nodefactors.colptr[K + 1]; k++) {
                node.index = K;
                node.region = s.nodefactors.region[k];
                node.fac = s.nodefactors.fac[k];
                node.x = &g.coord[(size_t)K * g.dim];
                for (int i = 0; i < NS; i++) {
                    u[i] = D(0.0);
                    u[i].d[i] = 1.0;
                    st[i] = D(0.0);
                }
                eval_storage(s.ph.slot[VFVM_SLOT_STORAGE], NS, st, u, node);
                const uint8_t* rs = &s.region_species[(size_t)(node.region - 1) * NS];
                for (int i = 0; i < NS; i++)
                    for (int j = 0; j < NS; j++)
                        if (rs[i] && rs[j]) out[((size_t)K * NS + i) * NS + j] += st[i].d[j] * node.fac;
            }
    };
    switch (n) {
        case 1: run(std::integral_constant<int, 1>()); break;
        case 2: run(std::integral_constant<int, 2>()); break;
        case 3: run(std::integral_constant<int, 3>()); break;
        case 4: run(std::integral_constant<int, 4>()); break;
        case 5: run(std::integral_constant<int, 5>()); break;
        case 10: run(std::integral_constant<int, 10>()); break;
        default: return VFVM_ERR_UNSUPPORTED;
    }
    return 0;
}

void vo_fbernoulli_pm(int n, const double* x, double* bp, double* bm, double* b) {
    for (int i = 0; i < n; i++) {
        fbernoulli_pm<double>(x[i], bp[i], bm[i]);
        b[i] = fbernoulli<double>(x[i]);
    }
}
// derivative of B(x) through the dual-number path (what ForwardDiff sees)
void vo_fbernoulli_dual(int n, const double* x, double* bp, double* dbp, double* bm, double* dbm) {
    for (int i = 0; i < n; i++) {
        Dual<1> xx(x[i]), p, m;
        xx.d[0] = 1.0;
        fbernoulli_pm(xx, p, m);
        bp[i] = p.v;
        dbp[i] = p.d[0];
        bm[i] = m.v;
        dbm[i] = m.d[0];
    }
}
void vo_cellfactors(int dim, int coordsys, const double* coord, const int* nodes, double* npar, double* epar) { cellfactors(dim, coordsys, coord, nodes, npar, epar); }
void vo_bfacefactors(int dim, int coordsys, const double* coord, const int* nodes, double* npar, double* epar) { bfacefactors(dim, coordsys, coord, nodes, npar, epar); }
int vo_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

// ---- CPU Krylov stand-in for `_solve_linear!` at sizes where a sparse LU is out of reach (src/vfvm_linsolve.jl:6-61 with
// method_linear = KrylovJL_CG / KrylovJL_BICGSTAB, precs = JacobiPreconBuilder / node-block BlockPreconBuilder) ---------------------
// Works on the matrix of the last vo_assemble.  The CSC matrix is transposed once into CSR with 32-bit column indices; SpMV, dots
// and vector updates run on `nthreads` OpenMP threads.  method: 0 BiCGStab, 1 CG; precon: 1 point Jacobi, 2 node-block (n x n) Jacobi.
#define VFVM_MAX_SPECIES_ORACLE 50
extern "C" int vo_krylov_solve(void* hv, int method, int precon, const double* b, double* x, double reltol, int maxit, int nthreads, int* iters_out, double* relres_out,
                               double* seconds_out) {
    System& s = *(System*)hv;
    const ExtMatrix& A = s.A;
    const int64_t nd = A.n;
    const int n = s.n;
    const int64_t nnz = (int64_t)A.nzval.size();
    if (nd <= 0 || nnz == 0) return VFVM_ERR_STATE;
    nthreads = std::max(1, nthreads);
    // CSC -> CSR
    std::vector<int64_t> rp((size_t)nd + 1, 0);
    for (int64_t k = 0; k < nnz; k++) rp[(size_t)A.rowval[k] + 1]++;
    for (int64_t i = 0; i < nd; i++) rp[i + 1] += rp[i];
    std::vector<int> ci((size_t)nnz);
    std::vector<double> va((size_t)nnz);
    {
        std::vector<int64_t> fill(rp.begin(), rp.end() - 1);
        for (int64_t j = 0; j < nd; j++)
            for (int64_t k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
                const int64_t o = fill[A.rowval[k]]++;
                ci[o] = (int)j;
                va[o] = A.nzval[k];
            }
    }
    const double t_start = omp_get_wtime();
    // inverse (block) diagonal
    const int bs = precon == 2 ? n : 1;
    std::vector<double> Minv((size_t)nd * bs, 0.0);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t K = 0; K < nd / bs; K++) {
        double B[VFVM_MAX_SPECIES_ORACLE][VFVM_MAX_SPECIES_ORACLE], I[VFVM_MAX_SPECIES_ORACLE][VFVM_MAX_SPECIES_ORACLE];
        for (int i = 0; i < bs; i++)
            for (int j = 0; j < bs; j++) {
                B[i][j] = 0.0;
                I[i][j] = i == j ? 1.0 : 0.0;
            }
        for (int i = 0; i < bs; i++)
            for (int64_t k = rp[K * bs + i]; k < rp[K * bs + i + 1]; k++) {
                const int64_t c = ci[k] - K * bs;
                if (c >= 0 && c < bs) B[i][c] = va[k];
            }
        for (int c = 0; c < bs; c++) {  // Gauss-Jordan without pivoting, as the device's node-block Jacobi
            const double ip = 1.0 / B[c][c];
            for (int j = 0; j < bs; j++) {
                B[c][j] *= ip;
                I[c][j] *= ip;
            }
            for (int i = 0; i < bs; i++) {
                if (i == c) continue;
                const double f = B[i][c];
                for (int j = 0; j < bs; j++) {
                    B[i][j] -= f * B[c][j];
                    I[i][j] -= f * I[c][j];
                }
            }
        }
        for (int i = 0; i < bs; i++)
            for (int j = 0; j < bs; j++) Minv[(size_t)(K * bs + i) * bs + j] = I[i][j];
    }
    auto spmv = [&](const double* in, double* out) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t i = 0; i < nd; i++) {
            double acc = 0.0;
            for (int64_t k = rp[i]; k < rp[i + 1]; k++) acc += va[k] * in[ci[k]];
            out[i] = acc;
        }
    };
    auto prec = [&](const double* in, double* out) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t K = 0; K < nd / bs; K++)
            for (int i = 0; i < bs; i++) {
                double acc = 0.0;
                for (int j = 0; j < bs; j++) acc += Minv[(size_t)(K * bs + i) * bs + j] * in[K * bs + j];
                out[K * bs + i] = acc;
            }
    };
    auto dot = [&](const double* a, const double* c) {
        double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(nthreads)
        for (int64_t i = 0; i < nd; i++) acc += a[i] * c[i];
        return acc;
    };
    std::vector<double> r(b, b + nd), z((size_t)nd), p((size_t)nd), q((size_t)nd);
    std::fill(x, x + nd, 0.0);
    const double bb = dot(b, b);
    double rr = bb;
    int it = 0;
    const double tol2 = reltol * reltol * bb;
    if (method == 1) {
        prec(r.data(), z.data());
        p = z;
        double rz = dot(r.data(), z.data());
        while (rr > tol2 && it < maxit) {
            it++;
            spmv(p.data(), q.data());
            const double alpha = rz / dot(p.data(), q.data());
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t i = 0; i < nd; i++) {
                x[i] += alpha * p[i];
                r[i] -= alpha * q[i];
            }
            prec(r.data(), z.data());
            const double rz_new = dot(r.data(), z.data());
            rr = dot(r.data(), r.data());
            const double beta = rz_new / rz;
            rz = rz_new;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t i = 0; i < nd; i++) p[i] = z[i] + beta * p[i];
        }
    } else {
        std::vector<double> rhat(r), v((size_t)nd, 0.0), sv((size_t)nd), t((size_t)nd), ph((size_t)nd), sh((size_t)nd);
        std::fill(p.begin(), p.end(), 0.0);
        double rho = 1.0, alpha = 1.0, omega = 1.0;
        while (rr > tol2 && it < maxit) {
            it++;
            const double rho_new = dot(rhat.data(), r.data());
            if (rho_new == 0.0 || omega == 0.0) break;
            const double beta = (rho_new / rho) * (alpha / omega);
            rho = rho_new;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t i = 0; i < nd; i++) p[i] = r[i] + beta * (p[i] - omega * v[i]);
            prec(p.data(), ph.data());
            spmv(ph.data(), v.data());
            alpha = rho / dot(rhat.data(), v.data());
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t i = 0; i < nd; i++) sv[i] = r[i] - alpha * v[i];
            prec(sv.data(), sh.data());
            spmv(sh.data(), t.data());
            const double tt = dot(t.data(), t.data());
            omega = tt > 0.0 ? dot(t.data(), sv.data()) / tt : 0.0;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t i = 0; i < nd; i++) {
                x[i] += alpha * ph[i] + omega * sh[i];
                r[i] = sv[i] - omega * t[i];
            }
            rr = dot(r.data(), r.data());
        }
    }
    // true residual
    spmv(x, q.data());
    double tr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tr) num_threads(nthreads)
    for (int64_t i = 0; i < nd; i++) tr += (b[i] - q[i]) * (b[i] - q[i]);
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = bb > 0.0 ? std::sqrt(tr / bb) : 0.0;
    if (seconds_out) *seconds_out = omp_get_wtime() - t_start;
    return 0;
}

// test/test040_inplacelu.jl probe: solves nsys systems with the oracle's restatement of inplace_linsolve!
extern "C" int vo_probe_inplace_linsolve(int n, int nsys, int pivoting, const double* A, const double* b, double* x) {
    std::vector<double> M((size_t)n * n), r(n);
    for (int s = 0; s < nsys; s++) {
        std::copy(A + (size_t)s * n * n, A + (size_t)(s + 1) * n * n, M.begin());
        std::copy(b + (size_t)s * n, b + (size_t)(s + 1) * n, r.begin());
        if (pivoting) inplace_linsolve_piv(n, M.data(), r.data());
        else inplace_linsolve_nopiv(n, M.data(), r.data());
        std::copy(r.begin(), r.end(), x + (size_t)s * n);
    }
    return 0;
}
