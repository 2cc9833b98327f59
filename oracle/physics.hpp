// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// CPU restatement of the physics callbacks the reference's examples define in Julia, written generically in
// the scalar type T (double or vo::Dual<P>) exactly like the Julia closures are generic in eltype(u).
// The ids / parameter layouts are the ones declared in include/vfvm_b200.h.
#pragma once
#include <cmath>
#include <vector>

#include "../include/vfvm_b200.h"
#include "dual.hpp"

namespace vo {

struct PhysSlot {
    int id = 0;
    std::vector<double> p;
};

struct BCEntry {
    int kind, species, region, has_ramp;
    double value, factor, t0, t1, v0, v1;
};

struct Physics {
    PhysSlot slot[VFVM_NUM_SLOTS];
    std::vector<BCEntry> bc;
    std::vector<double> nodal_source;  // n x N
};

// ---- geometry items handed to callbacks: src/vfvm_geometryitems.jl:102-184 (Node), :222-307 (BNode), :337-438 (Edge)
struct NodeCtx {
    int index = 0, region = 0, dim = 0;
    const double* x = nullptr;  // coordinates of the node
    double time = 0, embed = 0, fac = 0;
};
struct EdgeCtx {
    int index = 0, region = 0, dim = 0, nodeK = 0, nodeL = 0;
    const double *xK = nullptr, *xL = nullptr;
    double time = 0, embed = 0, fac = 0;
};
struct BNodeCtx {
    int index = 0, region = 0, dim = 0, ibface = 0, ibnode = 0;
    const double* x = nullptr;
    double time = 0, embed = 0, fac = 0;
    double Dirichlet = 1.0e30;          // src/vfvm_geometryitems.jl:296, src/vfvm_system.jl:329
    double* dirichlet_value = nullptr;  // n entries, src/vfvm_geometryitems.jl:259
};

// ---- src/vfvm_functions.jl:2-27 Horner scheme for B(x) around 0
template <class T>
inline T bernoulli_horner(const T& x) {
    const double c1 = 1.0 / 47900160.0, c2 = -1.0 / 1209600.0, c3 = 1.0 / 30240.0, c4 = -1.0 / 720.0, c5 = 1.0 / 12.0,
                 c6 = -1.0 / 2.0;
    T y = x * c1;
    y = x * y;
    y = x * (c2 + y);
    y = x * y;
    y = x * (c3 + y);
    y = x * y;
    y = x * (c4 + y);
    y = x * y;
    y = x * (c5 + y);
    y = x * (c6 + y);
    y = 1.0 + y;
    return y;
}

// ---- src/vfvm_functions.jl:49-59
template <class T>
inline T fbernoulli(const T& x) {
    const double small = 0.25, large = 50.0;  // :30-31
    double xv = value(x);
    if (xv < -large) return -x;
    if (xv > large) return T(0.0);
    if (std::fabs(xv) < small) return bernoulli_horner(x);
    return x / expm1(x);
}

// ---- src/vfvm_functions.jl:78-90: (B(x), B(-x))
template <class T>
inline void fbernoulli_pm(const T& x, T& bp, T& bm) {
    const double small = 0.25, large = 50.0;
    double xv = value(x);
    if (xv < -large) {
        bp = -x;
        bm = T(0.0);
    } else if (xv > large) {
        bp = T(0.0);
        bm = x;
    } else if (std::fabs(xv) < small) {
        T y = bernoulli_horner(x);
        bp = y;
        bm = x + y;
    } else {
        T y = x / expm1(x);
        bp = y;
        bm = x + y;
    }
}

// ---- src/vfvm_physics.jl:516-525
inline double ramp(double t, double tbegin, double tend, double ubegin, double uend) {
    if (t < tbegin) return ubegin;
    if (t < tend) return ubegin + (uend - ubegin) * (t - tbegin) / (tend - tbegin);
    return uend;
}

// ---------------------------------------------------------------- flux(f,u,edge,data)
// f must be zeroed by the caller (fwrap does y .= 0, src/vfvm_physics.jl:424-430)
// ---------------------------------------------------------------- inplace_linsolve!, src/vfvm_functions.jl:98-168
// non-pivoting Doolittle (doolittle_ludecomp! :98-115, doolittle_lusolve! :124-142); A row-major n x n, overwritten by L+U-I, b by x
template <class T>
inline void inplace_linsolve_nopiv(int n, T* A, T* b) {
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < i; j++) {
            for (int k = 0; k < j; k++) A[i * n + j] = A[i * n + j] - A[i * n + k] * A[k * n + j];
            A[i * n + j] = A[i * n + j] / A[j * n + j];
        }
        for (int j = i; j < n; j++)
            for (int k = 0; k < i; k++) A[i * n + j] = A[i * n + j] - A[i * n + k] * A[k * n + j];
    }
    for (int i = 0; i < n; i++)
        for (int k = 0; k < i; k++) b[i] = b[i] - A[i * n + k] * b[k];
    for (int i = n - 1; i >= 0; i--) {
        for (int k = i + 1; k < n; k++) b[i] = b[i] - A[i * n + k] * b[k];
        b[i] = b[i] / A[i * n + i];
    }
}
// LU with partial (row) pivoting on the VALUE of the entries, then the two triangular solves: what
// ldiv!(RecursiveFactorization.lu!(A, ipiv, Val(true), Val(false)), b) computes (:165-168)
template <class T>
inline void inplace_linsolve_piv(int n, T* A, T* b) {
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = std::fabs(value(A[c * n + c]));
        for (int r = c + 1; r < n; r++)
            if (std::fabs(value(A[r * n + c])) > best) {
                best = std::fabs(value(A[r * n + c]));
                piv = r;
            }
        if (piv != c) {
            for (int j = 0; j < n; j++) std::swap(A[c * n + j], A[piv * n + j]);
            std::swap(b[c], b[piv]);
        }
        for (int r = c + 1; r < n; r++) {
            A[r * n + c] = A[r * n + c] / A[c * n + c];
            for (int j = c + 1; j < n; j++) A[r * n + j] = A[r * n + j] - A[r * n + c] * A[c * n + j];
            b[r] = b[r] - A[r * n + c] * b[c];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        for (int k = i + 1; k < n; k++) b[i] = b[i] - A[i * n + k] * b[k];
        b[i] = b[i] / A[i * n + i];
    }
}

template <class T>
inline bool eval_flux(const PhysSlot& s, int n, T* f, const T* uK, const T* uL, const EdgeCtx& e) {
    const double* p = s.p.data();
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_FLUX_DIFFUSION:
            for (int i = 0; i < n; i++) f[i] = p[i] * (uK[i] - uL[i]);
            return true;
        case VFVM_FLUX_POWDIFF: {
            double m = p[n];
            for (int i = 0; i < n; i++) f[i] = p[i] * (powr(uK[i], m) - powr(uL[i], m));
            return true;
        }
        case VFVM_FLUX_MIXTURE: {  // examples/DevEx005_Mixture.jl:74-104
            std::vector<T> M((size_t)n * n, T(0.0)), au(n), du(n);
            for (int i = 0; i < n; i++) {
                M[(size_t)i * n + i] = T(1.0 / p[i]);
                du[i] = uK[i] - uL[i];
                au[i] = 0.5 * (uK[i] + uL[i]);
            }
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++)
                    if (i != j) {
                        M[(size_t)i * n + i] = M[(size_t)i * n + i] + au[j] / p[n + i * n + j];
                        M[(size_t)i * n + j] = -au[i] / p[n + i * n + j];
                    }
            inplace_linsolve_piv(n, M.data(), du.data());
            for (int i = 0; i < n; i++) f[i] = du[i];
            return true;
        }
        case VFVM_FLUX_CROSSDIFF2:
            f[0] = p[0] * (uK[0] - uL[0]) * (p[2] + uK[1] + uL[1]);
            f[1] = p[1] * (uK[1] - uL[1]) * (p[2] + uK[0] + uL[0]);
            return true;
        case VFVM_FLUX_SG_UNIPOLAR: {  // Example160 classflux!
            double eps = p[0];
            int iphi = (int)p[1], ic = (int)p[2];
            f[iphi] = eps * (uK[iphi] - uL[iphi]);
            T bp, bm;
            fbernoulli_pm(uK[iphi] - uL[iphi], bp, bm);
            f[ic] = bm * uK[ic] - bp * uL[ic];
            return true;
        }
        case VFVM_FLUX_SEDAN: {  // Example160 sedanflux!
            double eps = p[0], z = p[1];
            int iphi = (int)p[2], ic = (int)p[3];
            double eps_reg = p[4];
            f[iphi] = eps * (uK[iphi] - uL[iphi]);
            T mu1 = -log1p(maxr(-1.0 + eps_reg, -uK[ic]));
            T mu2 = -log1p(maxr(-1.0 + eps_reg, -uL[ic]));
            T bp, bm;
            fbernoulli_pm(z * 2.0 * (uK[iphi] - uL[iphi]) + (mu1 - mu2), bp, bm);
            f[ic] = bm * uK[ic] - bp * uL[ic];
            return true;
        }
        case VFVM_FLUX_SG_BIPOLAR: {  // Example161 flux!
            double lambda = p[0], mun = p[1], mup = p[2], zn = p[3], zp = p[4], En = p[5], Ep = p[6];
            int iphin = (int)p[7], iphip = (int)p[8], ipsi = (int)p[9];
            f[ipsi] = -(lambda * lambda) * (uL[ipsi] - uK[ipsi]);
            T bp, bm;
            fbernoulli_pm(-(uL[ipsi] - uK[ipsi]), bp, bm);
            T nn1 = exp(zn * (uK[iphin] - uK[ipsi] + En));
            T np1 = exp(zp * (uK[iphip] - uK[ipsi] + Ep));
            T nn2 = exp(zn * (uL[iphin] - uL[ipsi] + En));
            T np2 = exp(zp * (uL[iphip] - uL[ipsi] + Ep));
            f[iphin] = (-zn * mun) * (bm * nn2 - bp * nn1);
            f[iphip] = (-zp * mup) * (bp * np2 - bm * np1);
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- reaction(f,u,node,data)
template <class T>
inline bool eval_reaction(const PhysSlot& s, int n, T* f, const T* u, const NodeCtx& node) {
    const double* p = s.p.data();
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_REACTION_POW:
            for (int i = 0; i < n; i++) f[i] = p[i] * powr(u[i], p[n + i]);
            return true;
        case VFVM_REACTION_SINH:
            for (int i = 0; i < n; i++) f[i] = p[i] * (exp(u[i]) - exp(-u[i]));
            return true;
        case VFVM_REACTION_AFFINE:
            for (int i = 0; i < n; i++) {
                T acc(p[n * n + i]);
                for (int j = 0; j < n; j++)
                    if (p[i * n + j] != 0.0) acc = acc + p[i * n + j] * u[j];
                f[i] = acc;
            }
            return true;
        case VFVM_REACTION_REGION_AFFINE: {  // Example221 reaction: one affine map per cell region
            const int nreg = (int)p[0];
            if (node.region >= 1 && node.region <= nreg) {
                const double* q = p + 1 + (node.region - 1) * (n * n + n);
                for (int i = 0; i < n; i++) {
                    T acc(q[n * n + i]);
                    for (int j = 0; j < n; j++)
                        if (q[i * n + j] != 0.0) acc = acc + q[i * n + j] * u[j];
                    f[i] = acc;
                }
            }
            return true;
        }
        case VFVM_REACTION_BILINEAR2:
            f[0] = p[0] * (u[0] * u[1]);
            f[1] = (-p[0]) * (u[0] * u[1]);
            return true;
        case VFVM_REACTION_BIPOLAR: {  // Example161 reaction!
            double zn = p[0], zp = p[1], En = p[2], Ep = p[3], r0 = p[4];
            int iphin = (int)p[5], iphip = (int)p[6], ipsi = (int)p[7];
            int nreg = (int)p[8];
            double C = (node.region >= 1 && node.region <= nreg) ? p[9 + node.region - 1] : 0.0;
            T nn = exp(zn * (u[iphin] - u[ipsi] + En));
            T np = exp(zp * (u[iphip] - u[ipsi] + Ep));
            f[ipsi] = -(C + zn * nn + zp * np);
            T recomb = (r0 + 1.0 / (nn + np)) * (nn * np * (1.0 - exp(u[iphin] - u[iphip])));
            f[iphin] = zn * recomb;
            f[iphip] = zp * recomb;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- storage(f,u,node,data)
template <class T>
inline bool eval_storage(const PhysSlot& s, int n, T* f, const T* u, const NodeCtx& node) {
    const double* p = s.p.data();
    (void)node;
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_STORAGE_LINEAR:
            for (int i = 0; i < n; i++) f[i] = p[i] * u[i];
            return true;
        case VFVM_STORAGE_POW:
            for (int i = 0; i < n; i++) f[i] = powr(p[i] + u[i], 1.0 / p[n + i]);
            return true;
        case VFVM_STORAGE_BIPOLAR: {  // Example161 storage!
            double zn = p[0], zp = p[1], En = p[2], Ep = p[3];
            int iphin = (int)p[4], iphip = (int)p[5], ipsi = (int)p[6];
            T nn = exp(zn * (u[iphin] - u[ipsi] + En));
            T np = exp(zp * (u[iphip] - u[ipsi] + Ep));
            f[iphin] = zn * nn;
            f[iphip] = zp * np;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- source(f,node,data)
inline bool eval_source(const PhysSlot& s, int n, double* f, const NodeCtx& node, const std::vector<double>& nodal) {
    const double* p = s.p.data();
    const double* x = node.x;
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_SOURCE_CONST:
            for (int i = 0; i < n; i++) f[i] = p[i];
            return true;
        case VFVM_SOURCE_GAUSS: {
            int sp = (int)p[0];
            double a = p[1], r2 = 0.0;
            for (int d = 0; d < node.dim; d++) {
                double xd = x[d] - p[2 + d];
                r2 += xd * xd;
            }
            f[sp] = std::exp(-a * r2);
            return true;
        }
        case VFVM_SOURCE_XSINYEXPZ: {
            int sp = (int)p[0];
            f[sp] = x[0] * std::sin(p[1] * x[1]) * std::exp(x[2]);
            return true;
        }
        case VFVM_SOURCE_STEP1D: {
            int sp = (int)p[0];
            f[sp] = (x[0] <= p[1]) ? p[2] : p[3];
            return true;
        }
        case VFVM_SOURCE_AFFINE_X:
            for (int i = 0; i < n; i++) f[i] = p[i] + p[n + i] * x[0];
            return true;
        case VFVM_SOURCE_NODAL:
            for (int i = 0; i < n; i++) f[i] = nodal[(size_t)node.index * n + i];
            return true;
    }
    return false;
}

// ---------------------------------------------------------------- edgereaction(f,u,edge,data)
template <class T>
inline bool eval_edgereaction(const PhysSlot& s, int n, T* f, const T* uK, const T* uL, const EdgeCtx& edge) {
    const double* p = s.p.data();
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_EDGEREACTION_DIAMOND: {  // examples/DevEx002_EdgeReaction.jl:83-87: y = c h^2 / (2 dim), h = meas(edge)
            double h2 = 0.0;
            for (int d = 0; d < edge.dim; d++) h2 += (edge.xK[d] - edge.xL[d]) * (edge.xK[d] - edge.xL[d]);
            const double h = std::sqrt(h2);
            for (int i = 0; i < n; i++) f[i] = T(p[i] * (h * h) / (2 * edge.dim));
            return true;
        }
        case VFVM_EDGEREACTION_JOULE: {  // examples/Example206_JouleHeat.jl:83-86
            const int iphi = (int)p[1], iT = (int)p[2];
            f[iT] = -p[0] * (uK[iphi] - uL[iphi]) * (uK[iphi] - uL[iphi]);
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- bstorage(f,u,bnode,data)
template <class T>
inline bool eval_bstorage(const PhysSlot& s, int n, T* f, const T* u, const BNodeCtx& b) {
    const double* p = s.p.data();
    switch (s.id) {
        case VFVM_NONE: return true;
        case VFVM_BSTORAGE_LINEAR:  // examples/Example115_HeterogeneousCatalysis1D.jl:138-143
            if (b.region == (int)p[0])
                for (int i = 0; i < n; i++)
                    if (p[1 + i] != 0.0) f[i] = p[1 + i] * u[i];
            return true;
    }
    return false;
}

// ---------------------------------------------------------------- breaction(f,u,bnode,data) + BC helper calls
template <class T>
inline bool eval_breaction(const Physics& ph, int n, T* f, const T* u, BNodeCtx& b) {
    const PhysSlot& s = ph.slot[VFVM_SLOT_BREACTION];
    const double* p = s.p.data();
    switch (s.id) {
        case VFVM_NONE: break;
        case VFVM_BREACTION_LINEAR: {
            int r = (int)p[0];
            if (b.region == r) {
                for (int i = 0; i < n; i++) {
                    T acc(0.0);
                    for (int j = 0; j < n; j++)
                        if (p[1 + i * n + j] != 0.0) acc = acc + p[1 + i * n + j] * u[j];
                    f[i] = acc;
                }
            }
            break;
        }
        case VFVM_BREACTION_POW:  // examples/Example226_BoundaryIntegral.jl:42-47
            if (b.region == (int)p[0])
                for (int i = 0; i < n; i++)
                    if (p[1 + i] != 0.0) f[i] = p[1 + i] * powr(u[i], p[1 + n + i]);
            break;
        case VFVM_BREACTION_CATALYSIS: {  // examples/Example115_HeterogeneousCatalysis1D.jl:125-135
            if (b.region == (int)p[0]) {
                const double S = p[1], kpAC = p[2], kmAC = p[3], kpBC = p[4], kmBC = p[5];
                const int iA = (int)p[6], iB = (int)p[7], iC = (int)p[8];
                T rac = kpAC * u[iA] * (1.0 - u[iC]) - kmAC * u[iC];
                T rbc = kpBC * u[iB] * (1.0 - u[iC]) - kmBC * u[iC];
                f[iA] = S * rac;
                f[iB] = S * rbc;
                f[iC] = -rbc - rac;
            }
            break;
        }
        default: return false;
    }
    for (const BCEntry& e : ph.bc) {
        int ireg = e.region == 0 ? b.region : e.region;
        if (b.region != ireg) continue;
        double val = e.has_ramp ? ramp(b.time, e.t0, e.t1, e.v0, e.v1) : e.value;
        int i = e.species;
        switch (e.kind) {
            case VFVM_BC_DIRICHLET:  // src/vfvm_physics.jl:487-494
                f[i] = f[i] + b.Dirichlet * (u[i] - val);
                if (b.dirichlet_value) b.dirichlet_value[i] = val;
                break;
            case VFVM_BC_NEUMANN:  // :533
                f[i] = f[i] - val;
                break;
            case VFVM_BC_ROBIN:  // :552
                f[i] = f[i] + (e.factor * u[i] - val);
                break;
            default: return false;
        }
    }
    return true;
}

}  // namespace vo
