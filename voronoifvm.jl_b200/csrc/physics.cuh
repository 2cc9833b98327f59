// Registered device physics library (north_star item 2): every callback the kernels can evaluate.
// Each function is generic in the scalar type T (double or Dual<P>), like the Julia callbacks are generic in
// eltype(u); the reference lines each one restates are cited in include/vfvm_b200.h.
// FLUX is a template parameter of the row-tile kernel (it dominates register use); the node callbacks are
// selected by a warp-uniform runtime switch.
#pragma once
#include "dual.cuh"
#include "vfvm_internal.h"

// ---- Bernoulli function, src/vfvm_functions.jl:2-90 ---------------------------------------------
template <class T>
__device__ __forceinline__ T bernoulli_horner(const T& x) {
    const double c1 = 1.0 / 47900160.0, c2 = -1.0 / 1209600.0, c3 = 1.0 / 30240.0, c4 = -1.0 / 720.0, c5 = 1.0 / 12.0, c6 = -0.5;
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
    return 1.0 + y;
}

// (B(x), B(-x)); branches exactly as the reference: |x| < 0.25 Horner, |x| > 50 asymptotes, else x/expm1(x)
template <class T>
__device__ __forceinline__ void fbernoulli_pm(const T& x, T& bp, T& bm) {
    const double xv = dvalue(x);
    if (xv < -50.0) {
        bp = -x;
        bm = T(0.0);
    } else if (xv > 50.0) {
        bp = T(0.0);
        bm = x;
    } else if (fabs(xv) < 0.25) {
        T y = bernoulli_horner(x);
        bp = y;
        bm = x + y;
    } else {
        T y = x / dexpm1(x);
        bp = y;
        bm = x + y;
    }
}

__device__ __forceinline__ double ramp_fn(double t, double tbegin, double tend, double ubegin, double uend) {
    if (t < tbegin) return ubegin;
    if (t < tend) return ubegin + (uend - ubegin) * (t - tbegin) / (tend - tbegin);
    return uend;
}

// ---- inplace_linsolve!, src/vfvm_functions.jl:98-168: small dense systems solved inside a callback (DevEx005 mixture flux) ----------------
// non-pivoting Doolittle (doolittle_ludecomp! :98-115 + doolittle_lusolve! :124-142); A row-major, overwritten by L+U-I, b by the solution
template <int N, class T>
__device__ __forceinline__ void inplace_linsolve_nopiv(T* A, T* b) {
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < i; j++) {
            for (int k = 0; k < j; k++) A[i * N + j] = A[i * N + j] - A[i * N + k] * A[k * N + j];
            A[i * N + j] = A[i * N + j] / A[j * N + j];
        }
        for (int j = i; j < N; j++)
            for (int k = 0; k < i; k++) A[i * N + j] = A[i * N + j] - A[i * N + k] * A[k * N + j];
    }
    for (int i = 0; i < N; i++)
        for (int k = 0; k < i; k++) b[i] = b[i] - A[i * N + k] * b[k];
    for (int i = N - 1; i >= 0; i--) {
        for (int k = i + 1; k < N; k++) b[i] = b[i] - A[i * N + k] * b[k];
        b[i] = b[i] / A[i * N + i];
    }
}
// LU with partial (row) pivoting on the value of the entries + triangular solves: inplace_linsolve!(A, b, ipiv) (:165-168)
template <int N, class T>
__device__ __forceinline__ void inplace_linsolve_piv(T* A, T* b) {
    for (int c = 0; c < N; c++) {
        int piv = c;
        double best = fabs(dvalue(A[c * N + c]));
        for (int r = c + 1; r < N; r++) {
            const double v = fabs(dvalue(A[r * N + c]));
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (piv != c) {
            for (int j = 0; j < N; j++) {
                const T t = A[c * N + j];
                A[c * N + j] = A[piv * N + j];
                A[piv * N + j] = t;
            }
            const T t = b[c];
            b[c] = b[piv];
            b[piv] = t;
        }
        for (int r = c + 1; r < N; r++) {
            A[r * N + c] = A[r * N + c] / A[c * N + c];
            for (int j = c + 1; j < N; j++) A[r * N + j] = A[r * N + j] - A[r * N + c] * A[c * N + j];
            b[r] = b[r] - A[r * N + c] * b[c];
        }
    }
    for (int i = N - 1; i >= 0; i--) {
        for (int k = i + 1; k < N; k++) b[i] = b[i] - A[i * N + k] * b[k];
        b[i] = b[i] / A[i * N + i];
    }
}

// ---- flux(f,u,edge,data): f must be pre-zeroed ------------------------------------------------------
template <int FLUX, int NS, class T>
__device__ __forceinline__ void eval_flux(const double* __restrict__ p, T* f, const T* uK, const T* uL) {
    if constexpr (FLUX == VFVM_FLUX_DIFFUSION) {
#pragma unroll
        for (int i = 0; i < NS; i++) f[i] = p[i] * (uK[i] - uL[i]);
    } else if constexpr (FLUX == VFVM_FLUX_POWDIFF) {
        const double m = p[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) f[i] = p[i] * (dpowr(uK[i], m) - dpowr(uL[i], m));
    } else if constexpr (FLUX == VFVM_FLUX_CROSSDIFF2 && NS >= 2) {
        f[0] = p[0] * (uK[0] - uL[0]) * (p[2] + uK[1] + uL[1]);
        f[1] = p[1] * (uK[1] - uL[1]) * (p[2] + uK[0] + uL[0]);
    } else if constexpr (FLUX == VFVM_FLUX_SG_UNIPOLAR && NS == 2) {
        const double eps = p[0];
        const bool sw = ((int)p[1]) == 1;  // species order (iphi, ic) = (0,1) or (1,0); resolved without dynamic indexing
        const T &phiK = sw ? uK[1] : uK[0], &phiL = sw ? uL[1] : uL[0], &cK = sw ? uK[0] : uK[1], &cL = sw ? uL[0] : uL[1];
        T fphi = eps * (phiK - phiL);
        T bp, bm;
        fbernoulli_pm(phiK - phiL, bp, bm);
        T fc = bm * cK - bp * cL;
        f[0] = sw ? fc : fphi;
        f[1] = sw ? fphi : fc;
    } else if constexpr (FLUX == VFVM_FLUX_SEDAN && NS == 2) {
        const double eps = p[0], z = p[1], eps_reg = p[4];
        const bool sw = ((int)p[2]) == 1;
        const T &phiK = sw ? uK[1] : uK[0], &phiL = sw ? uL[1] : uL[0], &cK = sw ? uK[0] : uK[1], &cL = sw ? uL[0] : uL[1];
        T fphi = eps * (phiK - phiL);
        T mu1 = -dlog1p(dmaxr(-1.0 + eps_reg, -cK));
        T mu2 = -dlog1p(dmaxr(-1.0 + eps_reg, -cL));
        T bp, bm;
        fbernoulli_pm(z * 2.0 * (phiK - phiL) + (mu1 - mu2), bp, bm);
        T fc = bm * cK - bp * cL;
        f[0] = sw ? fc : fphi;
        f[1] = sw ? fphi : fc;
    } else if constexpr (FLUX == VFVM_FLUX_SG_BIPOLAR && NS == 3) {
        // species order fixed to (iphin, iphip, ipsi) = (0,1,2) on the device (checked at registration)
        const double lambda = p[0], mun = p[1], mup = p[2], zn = p[3], zp = p[4], En = p[5], Ep = p[6];
        f[2] = -(lambda * lambda) * (uL[2] - uK[2]);
        T bp, bm;
        fbernoulli_pm(-(uL[2] - uK[2]), bp, bm);
        T nn1 = dexp(zn * (uK[0] - uK[2] + En));
        T np1 = dexp(zp * (uK[1] - uK[2] + Ep));
        T nn2 = dexp(zn * (uL[0] - uL[2] + En));
        T np2 = dexp(zp * (uL[1] - uL[2] + Ep));
        f[0] = (-zn * mun) * (bm * nn2 - bp * nn1);
        f[1] = (-zp * mup) * (bp * np2 - bm * np1);
    } else if constexpr (FLUX == VFVM_FLUX_MIXTURE) {  // examples/DevEx005_Mixture.jl:74-104
        T M[NS * NS], au[NS], du[NS];
        for (int i = 0; i < NS; i++) {
            for (int j = 0; j < NS; j++) M[i * NS + j] = T(0.0);
            M[i * NS + i] = T(1.0 / p[i]);
            du[i] = uK[i] - uL[i];
            au[i] = 0.5 * (uK[i] + uL[i]);
        }
        for (int i = 0; i < NS; i++)
            for (int j = 0; j < NS; j++)
                if (i != j) {
                    M[i * NS + i] = M[i * NS + i] + au[j] / p[NS + i * NS + j];
                    M[i * NS + j] = -au[i] / p[NS + i * NS + j];
                }
        inplace_linsolve_piv<NS>(M, du);
        for (int i = 0; i < NS; i++) f[i] = du[i];
    }
}

// which (NS, FLUX) pairs have a device instantiation
__host__ __device__ constexpr bool flux_supported(int flux, int ns) {
    return flux == VFVM_NONE || flux == VFVM_FLUX_DIFFUSION || flux == VFVM_FLUX_POWDIFF || (flux == VFVM_FLUX_CROSSDIFF2 && ns == 2) ||
           (flux == VFVM_FLUX_SG_UNIPOLAR && ns == 2) || (flux == VFVM_FLUX_SEDAN && ns == 2) || (flux == VFVM_FLUX_SG_BIPOLAR && ns == 3) ||
           (flux == VFVM_FLUX_MIXTURE && (ns == 2 || ns == 3 || ns == 5));
}

// ---- reaction(f,u,node,data) -------------------------------------------------------------------------
template <int NS, class T>
__device__ __forceinline__ void eval_reaction(int id, const double* __restrict__ p, T* f, const T* u, int region) {
    switch (id) {
        case VFVM_REACTION_POW:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = p[i] * dpowr(u[i], p[NS + i]);
            break;
        case VFVM_REACTION_SINH:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = p[i] * (dexp(u[i]) - dexp(-u[i]));
            break;
        case VFVM_REACTION_AFFINE:
#pragma unroll
            for (int i = 0; i < NS; i++) {
                T acc(p[NS * NS + i]);
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    const double r = p[i * NS + j];
                    if (r != 0.0) acc = acc + r * u[j];
                }
                f[i] = acc;
            }
            break;
        case VFVM_REACTION_REGION_AFFINE: {
            const int nreg = (int)p[0];
            if (region >= 1 && region <= nreg) {
                const double* q = p + 1 + (region - 1) * (NS * NS + NS);
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    T acc(q[NS * NS + i]);
#pragma unroll
                    for (int j = 0; j < NS; j++) {
                        const double r = q[i * NS + j];
                        if (r != 0.0) acc = acc + r * u[j];
                    }
                    f[i] = acc;
                }
            }
            break;
        }
        case VFVM_REACTION_BILINEAR2:
            if constexpr (NS == 2) {
                f[0] = p[0] * (u[0] * u[1]);
                f[1] = (-p[0]) * (u[0] * u[1]);
            }
            break;
        case VFVM_REACTION_BIPOLAR:
            if constexpr (NS == 3) {
                const double zn = p[0], zp = p[1], En = p[2], Ep = p[3], r0 = p[4];
                const int nreg = (int)p[8];
                const double C = (region >= 1 && region <= nreg) ? p[9 + region - 1] : 0.0;
                T nn = dexp(zn * (u[0] - u[2] + En));
                T np = dexp(zp * (u[1] - u[2] + Ep));
                f[2] = -(C + zn * nn + zp * np);
                T recomb = (r0 + 1.0 / (nn + np)) * (nn * np * (1.0 - dexp(u[0] - u[1])));
                f[0] = zn * recomb;
                f[1] = zp * recomb;
            }
            break;
        default: break;
    }
}

// ---- storage(f,u,node,data) --------------------------------------------------------------------------
template <int NS, class T>
__device__ __forceinline__ void eval_storage(int id, const double* __restrict__ p, T* f, const T* u) {
    switch (id) {
        case VFVM_STORAGE_LINEAR:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = p[i] * u[i];
            break;
        case VFVM_STORAGE_POW:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = dpowr(p[i] + u[i], 1.0 / p[NS + i]);
            break;
        case VFVM_STORAGE_BIPOLAR:
            if constexpr (NS == 3) {
                const double zn = p[0], zp = p[1], En = p[2], Ep = p[3];
                T nn = dexp(zn * (u[0] - u[2] + En));
                T np = dexp(zp * (u[1] - u[2] + Ep));
                f[0] = zn * nn;
                f[1] = zp * np;
            }
            break;
        default: break;
    }
}

// ---- source(f,node,data) -----------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void eval_source(int id, const double* __restrict__ p, double* f, const double* __restrict__ x, int dim,
                                            const double* __restrict__ nodal, int64_t K) {
    switch (id) {
        case VFVM_SOURCE_CONST:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = p[i];
            break;
        case VFVM_SOURCE_GAUSS: {
            const int sp = (int)p[0];
            double r2 = 0.0;
            for (int d = 0; d < dim; d++) {
                const double xd = x[d] - p[2 + d];
                r2 += xd * xd;
            }
            const double s = exp(-p[1] * r2);
#pragma unroll
            for (int i = 0; i < NS; i++)
                if (i == sp) f[i] = s;
            break;
        }
        case VFVM_SOURCE_XSINYEXPZ: {
            const int sp = (int)p[0];
            const double s = x[0] * sin(p[1] * x[1]) * exp(x[2]);
#pragma unroll
            for (int i = 0; i < NS; i++)
                if (i == sp) f[i] = s;
            break;
        }
        case VFVM_SOURCE_STEP1D: {
            const int sp = (int)p[0];
            const double s = (x[0] <= p[1]) ? p[2] : p[3];
#pragma unroll
            for (int i = 0; i < NS; i++)
                if (i == sp) f[i] = s;
            break;
        }
        case VFVM_SOURCE_AFFINE_X:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = p[i] + p[NS + i] * x[0];
            break;
        case VFVM_SOURCE_NODAL:
#pragma unroll
            for (int i = 0; i < NS; i++) f[i] = nodal[K * NS + i];
            break;
        default: break;
    }
}

// ---- breaction(f,u,bnode,data) + boundary_dirichlet!/neumann!/robin! calls --------------------------
// dirichlet_value (may be null) receives the values set by boundary_dirichlet! (src/vfvm_physics.jl:492)
// the registered boundary reaction function alone (no boundary_dirichlet!/neumann!/robin! helper calls)
template <int NS, class T>
__device__ __forceinline__ void eval_breaction_fn(int id, const double* __restrict__ p, T* f, const T* u, int bregion) {
    if (id == VFVM_BREACTION_LINEAR) {
        if (bregion == (int)p[0]) {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                T acc(0.0);
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    const double r = p[1 + i * NS + j];
                    if (r != 0.0) acc = acc + r * u[j];
                }
                f[i] = acc;
            }
        }
    } else if (id == VFVM_BREACTION_POW) {  // examples/Example226_BoundaryIntegral.jl:42-47
        if (bregion == (int)p[0]) {
#pragma unroll
            for (int i = 0; i < NS; i++)
                if (p[1 + i] != 0.0) f[i] = p[1 + i] * dpowr(u[i], p[1 + NS + i]);
        }
    } else if (id == VFVM_BREACTION_CATALYSIS) {  // examples/Example115_HeterogeneousCatalysis1D.jl:125-135
        if constexpr (NS >= 3) {
            if (bregion == (int)p[0]) {
                const double S = p[1], kpAC = p[2], kmAC = p[3], kpBC = p[4], kmBC = p[5];
                const int iA = (int)p[6], iB = (int)p[7], iC = (int)p[8];
                T uA(0.0), uB(0.0), uC(0.0);
#pragma unroll
                for (int i = 0; i < NS; i++) {  // register arrays: select instead of indexing
                    if (i == iA) uA = u[i];
                    if (i == iB) uB = u[i];
                    if (i == iC) uC = u[i];
                }
                const T rac = kpAC * uA * (1.0 - uC) - kmAC * uC;
                const T rbc = kpBC * uB * (1.0 - uC) - kmBC * uC;
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    if (i == iA) f[i] = S * rac;
                    if (i == iB) f[i] = S * rbc;
                    if (i == iC) f[i] = -rbc - rac;
                }
            }
        }
    }
}

template <int NS, class T>
__device__ __forceinline__ void eval_breaction(const PhysicsDev& ph, T* f, const T* u, int bregion, double time, double Dirichlet,
                                               double* dirichlet_value) {
    const PhysSlotDev& s = ph.slot[VFVM_SLOT_BREACTION];
    eval_breaction_fn<NS>(s.id, ph.params + s.off, f, u, bregion);
    for (int e = 0; e < ph.nbc; e++) {
        const vfvm_bc_entry& bc = ph.bc[e];
        const int ireg = bc.region == 0 ? bregion : bc.region;
        if (bregion != ireg) continue;
        const double val = bc.has_ramp ? ramp_fn(time, bc.t0, bc.t1, bc.v0, bc.v1) : bc.value;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (i != bc.species) continue;
            if (bc.kind == VFVM_BC_DIRICHLET) {
                f[i] = f[i] + Dirichlet * (u[i] - val);
                if (dirichlet_value) dirichlet_value[i] = val;
            } else if (bc.kind == VFVM_BC_NEUMANN) {
                f[i] = f[i] - val;
            } else if (bc.kind == VFVM_BC_ROBIN) {
                f[i] = f[i] + (bc.factor * u[i] - val);
            }
        }
    }
}

// ---- bstorage(f,u,bnode,data): f pre-zeroed ---------------------------------------------------------------------------------------
template <int NS, class T>
__device__ __forceinline__ void eval_bstorage_fn(int id, const double* __restrict__ p, T* f, const T* u, int bregion) {
    if (id == VFVM_BSTORAGE_LINEAR && bregion == (int)p[0]) {  // examples/Example115_HeterogeneousCatalysis1D.jl:138-143
#pragma unroll
        for (int i = 0; i < NS; i++)
            if (p[1 + i] != 0.0) f[i] = p[1 + i] * u[i];
    }
}
template <int NS, class T>
__device__ __forceinline__ void eval_bstorage(const PhysicsDev& ph, T* f, const T* u, int bregion) {
    const PhysSlotDev& s = ph.slot[VFVM_SLOT_BSTORAGE];
    eval_bstorage_fn<NS>(s.id, ph.params + s.off, f, u, bregion);
}

// ---- edgereaction(f,u,edge,data): f pre-zeroed; h = meas(edge) ------------------------------------------------------------------
template <int NS, class T>
__device__ __forceinline__ void eval_edgereaction(int id, const double* __restrict__ p, T* f, const T* uK, const T* uL, double h, int dim) {
    if (id == VFVM_EDGEREACTION_DIAMOND) {  // examples/DevEx002_EdgeReaction.jl:83-87
#pragma unroll
        for (int i = 0; i < NS; i++) f[i] = T(p[i] * (h * h) / (2 * dim));
    } else if (id == VFVM_EDGEREACTION_JOULE) {  // examples/Example206_JouleHeat.jl:83-86
        const int iphi = (int)p[1], iT = (int)p[2];
        T d(0.0);
#pragma unroll
        for (int i = 0; i < NS; i++)
            if (i == iphi) d = uK[i] - uL[i];
#pragma unroll
        for (int i = 0; i < NS; i++)
            if (i == iT) f[i] = -p[0] * d * d;
    }
}
