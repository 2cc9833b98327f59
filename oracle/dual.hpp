// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// Forward-mode dual numbers with a full chunk of P partials, restating what ForwardDiff.jl does for
// VoronoiFVM's ResJacEvaluator (src/vfvm_physics.jl:437-449: chunk = length(input), jacobian!).
// ForwardDiff itself is a third-party dependency that is not vendored under /root/reference
// (Project.toml: ForwardDiff = "0.10.35, 1"); the rules below are its published differentiation rules
// (DiffRules): d(a*b) = a'b + ab', d(a/b) = (a' - (a/b) b')/b, d exp = exp, d expm1 = exp, d log1p = 1/(1+x),
// d x^p = p x^(p-1), literal x^2 = x*x.
#pragma once
#include <cmath>

namespace vo {

template <int P>
struct Dual {
    double v;
    double d[P];
    Dual() : v(0.0) {
        for (int i = 0; i < P; i++) d[i] = 0.0;
    }
    Dual(double x) : v(x) {
        for (int i = 0; i < P; i++) d[i] = 0.0;
    }
};

template <int P>
inline Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v + b.v;
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int P>
inline Dual<P> operator+(const Dual<P>& a, double b) {
    Dual<P> r = a;
    r.v = a.v + b;
    return r;
}
template <int P>
inline Dual<P> operator+(double a, const Dual<P>& b) {
    return b + a;
}
template <int P>
inline Dual<P> operator-(const Dual<P>& a) {
    Dual<P> r;
    r.v = -a.v;
    for (int i = 0; i < P; i++) r.d[i] = -a.d[i];
    return r;
}
template <int P>
inline Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v - b.v;
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int P>
inline Dual<P> operator-(const Dual<P>& a, double b) {
    Dual<P> r = a;
    r.v = a.v - b;
    return r;
}
template <int P>
inline Dual<P> operator-(double a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a - b.v;
    for (int i = 0; i < P; i++) r.d[i] = -b.d[i];
    return r;
}
template <int P>
inline Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v * b.v;
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int P>
inline Dual<P> operator*(const Dual<P>& a, double b) {
    Dual<P> r;
    r.v = a.v * b;
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b;
    return r;
}
template <int P>
inline Dual<P> operator*(double a, const Dual<P>& b) {
    return b * a;
}
template <int P>
inline Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a.v / b.v;
    for (int i = 0; i < P; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
template <int P>
inline Dual<P> operator/(const Dual<P>& a, double b) {
    Dual<P> r;
    r.v = a.v / b;
    for (int i = 0; i < P; i++) r.d[i] = a.d[i] / b;
    return r;
}
template <int P>
inline Dual<P> operator/(double a, const Dual<P>& b) {
    Dual<P> r;
    r.v = a / b.v;
    for (int i = 0; i < P; i++) r.d[i] = -(r.v * b.d[i]) / b.v;
    return r;
}

inline double value(double x) { return x; }
template <int P>
inline double value(const Dual<P>& x) {
    return x.v;
}

template <int P>
inline Dual<P> exp(const Dual<P>& a) {
    Dual<P> r;
    r.v = std::exp(a.v);
    for (int i = 0; i < P; i++) r.d[i] = r.v * a.d[i];
    return r;
}
template <int P>
inline Dual<P> expm1(const Dual<P>& a) {
    Dual<P> r;
    r.v = std::expm1(a.v);
    double e = std::exp(a.v);
    for (int i = 0; i < P; i++) r.d[i] = e * a.d[i];
    return r;
}
template <int P>
inline Dual<P> log1p(const Dual<P>& a) {
    Dual<P> r;
    r.v = std::log1p(a.v);
    double g = 1.0 / (1.0 + a.v);
    for (int i = 0; i < P; i++) r.d[i] = g * a.d[i];
    return r;
}
template <int P>
inline Dual<P> sqrt(const Dual<P>& a) {
    Dual<P> r;
    r.v = std::sqrt(a.v);
    double g = 0.5 / r.v;
    for (int i = 0; i < P; i++) r.d[i] = g * a.d[i];
    return r;
}
// x^p for a real exponent p; p == 2 follows Julia's literal_pow (x*x)
template <int P>
inline Dual<P> powr(const Dual<P>& a, double p) {
    if (p == 2.0) return a * a;
    if (p == 1.0) return a;
    Dual<P> r;
    r.v = std::pow(a.v, p);
    double g = p * std::pow(a.v, p - 1.0);
    for (int i = 0; i < P; i++) r.d[i] = g * a.d[i];
    return r;
}
// max(real, dual): ForwardDiff picks the operand with the larger value
template <int P>
inline Dual<P> maxr(double a, const Dual<P>& b) {
    return (b.v < a) ? Dual<P>(a) : b;
}

inline double exp(double x) { return std::exp(x); }
inline double expm1(double x) { return std::expm1(x); }
inline double log1p(double x) { return std::log1p(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double powr(double a, double p) {
    if (p == 2.0) return a * a;
    if (p == 1.0) return a;
    return std::pow(a, p);
}
inline double maxr(double a, double b) { return (b < a) ? a : b; }

}  // namespace vo
