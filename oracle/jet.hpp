// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under oracle/ may be imported, linked or executed by the product path.
//
// Jet2<N>: dense second-order forward-mode scalar (value, gradient, full symmetric Hessian)
// in N independents.  It plays the role CppAD's tape plays in the reference
// (lion::ipopt_cppad_solve, call site /root/reference/src/core/applications/optimal_laptime.hpp:546-551):
// exact first and second derivatives of the straight-line model code, up to rounding.
#ifndef ORACLE_JET_HPP
#define ORACLE_JET_HPP

#include <cmath>

// The real type the jets carry.  double: the oracle proper (what every test compares with).  Built a second time with
// -DORACLE_REAL="long double" (liboracle_ld.so: x87 extended precision, 64-bit mantissa) the same formulas give second
// derivatives with three more digits -- the arbiter where fp64 cancellation inside the tyre model limits the entry-wise
// agreement of two correct double evaluations to ~1e-8 (tests/test_gpu_parity.py).
#ifndef ORACLE_REAL
#define ORACLE_REAL double
#endif
typedef ORACLE_REAL oreal;

template <int N>
struct Jet2
{
    oreal v;
    oreal g[N];
    oreal h[N][N];   // full symmetric storage, only j<=i is maintained and read

    Jet2() : v(0.0)
    {
        for (int i = 0; i < N; ++i) { g[i] = 0.0; for (int j = 0; j <= i; ++j) h[i][j] = 0.0; }
    }
    Jet2(oreal c) : v(c)
    {
        for (int i = 0; i < N; ++i) { g[i] = 0.0; for (int j = 0; j <= i; ++j) h[i][j] = 0.0; }
    }
    static Jet2 variable(oreal value, int index)
    {
        Jet2 r(value);
        if (index >= 0 && index < N) r.g[index] = 1.0;
        return r;
    }
    oreal hess(int i, int j) const { return (j <= i) ? h[i][j] : h[j][i]; }
};

// y = f(a) given f, f', f'' at a.v
template <int N>
inline Jet2<N> jet_chain(const Jet2<N>& a, oreal f0, oreal f1, oreal f2)
{
    Jet2<N> r;
    r.v = f0;
    for (int i = 0; i < N; ++i) r.g[i] = f1 * a.g[i];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j)
            r.h[i][j] = f1 * a.h[i][j] + f2 * a.g[i] * a.g[j];
    return r;
}

template <int N> inline Jet2<N> operator+(const Jet2<N>& a, const Jet2<N>& b)
{
    Jet2<N> r;
    r.v = a.v + b.v;
    for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) r.h[i][j] = a.h[i][j] + b.h[i][j];
    return r;
}
template <int N> inline Jet2<N> operator-(const Jet2<N>& a, const Jet2<N>& b)
{
    Jet2<N> r;
    r.v = a.v - b.v;
    for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i];
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) r.h[i][j] = a.h[i][j] - b.h[i][j];
    return r;
}
template <int N> inline Jet2<N> operator-(const Jet2<N>& a)
{
    Jet2<N> r;
    r.v = -a.v;
    for (int i = 0; i < N; ++i) r.g[i] = -a.g[i];
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) r.h[i][j] = -a.h[i][j];
    return r;
}
template <int N> inline Jet2<N> operator*(const Jet2<N>& a, const Jet2<N>& b)
{
    Jet2<N> r;
    r.v = a.v * b.v;
    for (int i = 0; i < N; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j)
            r.h[i][j] = a.v * b.h[i][j] + b.v * a.h[i][j] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
    return r;
}
template <int N> inline Jet2<N> inv(const Jet2<N>& b)
{
    const oreal f0 = 1.0 / b.v;
    return jet_chain(b, f0, -f0 * f0, 2.0 * f0 * f0 * f0);
}
template <int N> inline Jet2<N> operator/(const Jet2<N>& a, const Jet2<N>& b) { return a * inv(b); }

// mixed with double
template <int N> inline Jet2<N> operator+(const Jet2<N>& a, double b) { Jet2<N> r = a; r.v += b; return r; }
template <int N> inline Jet2<N> operator+(double a, const Jet2<N>& b) { return b + a; }
template <int N> inline Jet2<N> operator-(const Jet2<N>& a, double b) { Jet2<N> r = a; r.v -= b; return r; }
template <int N> inline Jet2<N> operator-(double a, const Jet2<N>& b) { return (-b) + a; }
template <int N> inline Jet2<N> operator*(const Jet2<N>& a, double b)
{
    Jet2<N> r;
    r.v = a.v * b;
    for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b;
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) r.h[i][j] = a.h[i][j] * b;
    return r;
}
template <int N> inline Jet2<N> operator*(double a, const Jet2<N>& b) { return b * a; }
template <int N> inline Jet2<N> operator/(const Jet2<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Jet2<N> operator/(double a, const Jet2<N>& b) { return inv(b) * a; }
template <int N> inline Jet2<N>& operator+=(Jet2<N>& a, const Jet2<N>& b) { a = a + b; return a; }
template <int N> inline Jet2<N>& operator-=(Jet2<N>& a, const Jet2<N>& b) { a = a - b; return a; }
template <int N> inline Jet2<N>& operator+=(Jet2<N>& a, double b) { a.v += b; return a; }
template <int N> inline Jet2<N>& operator*=(Jet2<N>& a, const Jet2<N>& b) { a = a * b; return a; }
template <int N> inline Jet2<N>& operator*=(Jet2<N>& a, double b) { a = a * b; return a; }

template <int N> inline Jet2<N> sqrt(const Jet2<N>& a)
{
    const oreal s = std::sqrt(a.v);
    return jet_chain(a, s, 0.5 / s, -0.25 / (s * a.v));
}
template <int N> inline Jet2<N> sin(const Jet2<N>& a)
{
    const oreal s = std::sin(a.v), c = std::cos(a.v);
    return jet_chain(a, s, c, -s);
}
template <int N> inline Jet2<N> cos(const Jet2<N>& a)
{
    const oreal s = std::sin(a.v), c = std::cos(a.v);
    return jet_chain(a, c, -s, -c);
}
template <int N> inline Jet2<N> atan(const Jet2<N>& a)
{
    const oreal d = 1.0 / (1.0 + a.v * a.v);
    return jet_chain(a, std::atan(a.v), d, -2.0 * a.v * d * d);
}
template <int N> inline Jet2<N> tanh(const Jet2<N>& a)
{
    const oreal t = std::tanh(a.v);
    const oreal d = 1.0 - t * t;
    return jet_chain(a, t, d, -2.0 * t * d);
}
template <int N> inline Jet2<N> exp(const Jet2<N>& a)
{
    const oreal e = std::exp(a.v);
    return jet_chain(a, e, e, e);
}

inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Jet2<N>& x) { return (double)x.v; }

#endif
