// fp64 forward-mode dual numbers for device code: value + P partial derivatives held in registers.
// Replaces ForwardDiff.jl inside ResJacEvaluator (src/vfvm_physics.jl:394-466: full chunk, jacobian!).
// All loops have compile-time bounds so that ptxas keeps d[] in registers (never index d[] dynamically).
#pragma once
#include <cuda_runtime.h>

template <int P>
struct Dual {
    double v;
    double d[P];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < P; i++) d[i] = 0.0;
    }
};

#define DUAL_FN template <int P> __device__ __forceinline__

DUAL_FN Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
DUAL_FN Dual<P> operator+(const Dual<P>& a, double b) {
    Dual<P> r = a;
    r.v = a.v + b;
    return r;
}
DUAL_FN Dual<P> operator+(double a, const Dual<P>& b) { return b + a; }
DUAL_FN Dual<P> operator-(const Dual<P>& a) {
    Dual<P> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = -a.d[i];
    return r;
}
DUAL_FN Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
DUAL_FN Dual<P> operator-(const Dual<P>& a, double b) {
    Dual<P> r = a;
    r.v = a.v - b;
    return r;
}
DUAL_FN Dual<P> operator-(double a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a - b.v;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = -b.d[i];
    return r;
}
DUAL_FN Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
DUAL_FN Dual<P> operator*(const Dual<P>& a, double b) {
    Dual<P> r;
    r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b;
    return r;
}
DUAL_FN Dual<P> operator*(double a, const Dual<P>& b) { return b * a; }
DUAL_FN Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    double ib = 1.0 / b.v;
    r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
DUAL_FN Dual<P> operator/(const Dual<P>& a, double b) {
    Dual<P> r;
    r.v = a.v / b;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] / b;
    return r;
}
DUAL_FN Dual<P> operator/(double a, const Dual<P>& b) {
    Dual<P> r;
    double ib = 1.0 / b.v;
    r.v = a * ib;
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = -(r.v * b.d[i]) * ib;
    return r;
}

__device__ __forceinline__ double dvalue(double x) { return x; }
DUAL_FN double dvalue(const Dual<P>& x) { return x.v; }

DUAL_FN Dual<P> dexp(const Dual<P>& a) {
    Dual<P> r;
    r.v = exp(a.v);
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = r.v * a.d[i];
    return r;
}
DUAL_FN Dual<P> dexpm1(const Dual<P>& a) {
    Dual<P> r;
    r.v = expm1(a.v);
    double e = r.v + 1.0;  // exp(x) = expm1(x) + 1, exact to 1 ulp for |x| >= 0.25 (the only range this is used in)
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = e * a.d[i];
    return r;
}
DUAL_FN Dual<P> dlog1p(const Dual<P>& a) {
    Dual<P> r;
    r.v = log1p(a.v);
    double g = 1.0 / (1.0 + a.v);
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = g * a.d[i];
    return r;
}
// x^p, real exponent; p == 2 is Julia's literal_pow (x*x)
DUAL_FN Dual<P> dpowr(const Dual<P>& a, double p) {
    if (p == 2.0) return a * a;
    if (p == 1.0) return a;
    Dual<P> r;
    r.v = pow(a.v, p);
    double g = p * pow(a.v, p - 1.0);
#pragma unroll
    for (int i = 0; i < P; i++) r.d[i] = g * a.d[i];
    return r;
}
DUAL_FN Dual<P> dmaxr(double a, const Dual<P>& b) { return (b.v < a) ? Dual<P>(a) : b; }

__device__ __forceinline__ double dexp(double x) { return exp(x); }
__device__ __forceinline__ double dexpm1(double x) { return expm1(x); }
__device__ __forceinline__ double dlog1p(double x) { return log1p(x); }
__device__ __forceinline__ double dpowr(double a, double p) {
    if (p == 2.0) return a * a;
    if (p == 1.0) return a;
    return pow(a, p);
}
__device__ __forceinline__ double dmaxr(double a, double b) { return (b < a) ? a : b; }
