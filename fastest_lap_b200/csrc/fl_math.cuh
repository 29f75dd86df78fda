// Branch-free fp64 elementary functions for the generated node programs (values pass).
//
// The CUDA math library's sin / cos / atan / tanh / exp / sqrt / division each carry a rarely taken slow path behind a
// branch.  A node program calls ~70 of them; every branch ends a scheduling region, so the four tyres' independent chains
// are issued one after the other (profiles/prof_r1f_values: 344 static branches, 7.4 cycles per issued instruction, 23 % of
// the issued instructions are constant moves).  The versions below are straight-line code -- argument reduction by
// fused multiply-adds, polynomial kernels with the classical fdlibm (Sun Microsystems, 1993) minimax coefficients,
// selections instead of branches -- valid for finite arguments of moderate size (|x| < 1e5 for the trigonometric
// reduction), accurate to about 1 ulp (tests/test_flmath.py on the host, tests/test_gpu_parity.py on the device).
// Outside those ranges they return finite garbage or NaN, which the solver's line search rejects like any other
// non-finite trial point.  Seeds come from MUFU.RCP64H / MUFU.RSQ64H on the device and from float arithmetic on the host
// (the host build exists for the accuracy test only).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define FLM __host__ __device__ __forceinline__
#else
#define FLM inline
#endif

FLM double flm_rcp_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return (double)(1.0f / (float)x);
#endif
}

FLM double flm_rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return (double)(1.0f / sqrtf((float)x));
#endif
}

FLM int flm_rint(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2int_rn(x);
#else
    return (int)nearbyint(x);
#endif
}

FLM double flm_pow2(int k)          // 2^k, -1022 <= k <= 1023
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((k + 1023) << 20, 0);
#else
    const uint64_t b = (uint64_t)(k + 1023) << 52;
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}

// 1 / x: three Newton steps on the seed (the seed carries 20+ bits on the device, 24 on the host)
FLM double fl_rcp(double x)
{
    double y = flm_rcp_seed(x);
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// n / d with one residual correction
FLM double fl_div(double n, double d)
{
    const double y = fl_rcp(d);
    const double q = n * y;
    return fma(fma(-d, q, n), y, q);
}

FLM double fl_sqrt(double x)
{
    double r = flm_rsqrt_seed(x);
    const double h = 0.5 * x;
    r = r * fma(-h * r, r, 1.5);
    r = r * fma(-h * r, r, 1.5);
    r = r * fma(-h * r, r, 1.5);
    double s = x * r;
    s = fma(fma(-s, s, x), 0.5 * r, s);
    return x == 0.0 ? 0.0 : s;
}

// sin and cos of r in [-pi/4, pi/4]: fdlibm k_sin.c / k_cos.c
FLM void flm_sincos_kernel(double r, double& s, double& c)
{
    const double z = r * r;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08), 2.75573137070700676789e-06),
                                              -1.98412698298579493134e-04), 8.33333333332248946124e-03), -1.66666666666666324348e-01);
    s = fma(r * z, ps, r);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09), -2.75573143513906633035e-07),
                                              2.48015872894767294178e-05), -1.38888888888741095749e-03), 4.16666666666666019037e-02);
    c = fma(z * z, pc, fma(z, -0.5, 1.0));
}

FLM void fl_sincos(double x, double* sn, double* cs)
{
    const int q = flm_rint(x * 0.63661977236758134308);          // 2 / pi
    const double k = (double)q;
    double r = fma(-k, 1.57079632679489655800e+00, x);            // pi / 2 in two pieces
    r = fma(-k, 6.12323399573676603587e-17, r);
    double s, c;
    flm_sincos_kernel(r, s, c);
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

FLM double fl_sin(double x) { double s, c; fl_sincos(x, &s, &c); return s; }
FLM double fl_cos(double x) { double s, c; fl_sincos(x, &s, &c); return c; }

// fdlibm s_atan.c with the interval chosen by selections
FLM double fl_atan(double x)
{
    const double ax = fabs(x);
    // every interval's quotient operands are formed first, so that the choice is a chain of plain selections
    const double n1 = fma(2.0, ax, -1.0), d1 = 2.0 + ax, n2 = ax - 1.0, d2 = ax + 1.0, n3 = ax - 1.5, d3 = fma(1.5, ax, 1.0);
    double num = -1.0, den = ax, hi = 1.57079632679489655800e+00, lo = 6.12323399573676603587e-17;
    const bool r3 = ax < 2.4375, r2 = ax < 1.1875, r1 = ax < 0.6875, r0 = ax < 0.4375;
    num = r3 ? n3 : num; den = r3 ? d3 : den; hi = r3 ? 9.82793723247329054082e-01 : hi; lo = r3 ? 1.39033110312309984516e-17 : lo;
    num = r2 ? n2 : num; den = r2 ? d2 : den; hi = r2 ? 7.85398163397448278999e-01 : hi; lo = r2 ? 3.06161699786838301793e-17 : lo;
    num = r1 ? n1 : num; den = r1 ? d1 : den; hi = r1 ? 4.63647609000806093515e-01 : hi; lo = r1 ? 2.26987774529616870924e-17 : lo;
    num = r0 ? ax : num; den = r0 ? 1.0 : den; hi = r0 ? 0.0 : hi; lo = r0 ? 0.0 : lo;
    const double t = fl_div(num, den);
    const double z = t * t, w = z * z;
    const double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, 1.62858201153657823623e-02, 4.97687799461593236017e-02), 6.66107313738753120669e-02),
                                                9.09088713343650656196e-02), 1.42857142725034663711e-01), 3.33333333333329318027e-01);
    const double s2 = w * fma(w, fma(w, fma(w, fma(w, -3.65315727442169155270e-02, -5.83357013379057348645e-02), -7.69187620504482999495e-02),
                                         -1.11111104054623557880e-01), -1.99999999998764832476e-01);
    const double res = hi - ((t * (s1 + s2) - lo) - t);
    return x < 0.0 ? -res : res;
}

// exp(y) - 1 = 2^k (1 + p) - 1 with y = k ln 2 + r, p = exp(r) - 1 by its Taylor polynomial (|r| <= 0.347: the degree-13
// remainder is below 1e-17); y is clamped to [-700, 700]
FLM void flm_exp_parts(double y, double& scale, double& p)
{
    y = fmin(fmax(y, -700.0), 700.0);
    const int k = flm_rint(y * 1.44269504088896338700e+00);
    const double kd = (double)k;
    double r = fma(-kd, 6.93147180369123816490e-01, y);           // ln 2 in two pieces (fdlibm e_exp.c)
    r = fma(-kd, 1.90821492927058770002e-10, r);
    double q = 1.0 / 6227020800.0;
    q = fma(q, r, 1.0 / 479001600.0);
    q = fma(q, r, 1.0 / 39916800.0);
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    p = fma(r * r, q, r);
    scale = flm_pow2(k);
}

FLM double fl_exp(double y)
{
    double sc, p;
    flm_exp_parts(y, sc, p);
    return y != y ? y : fma(sc, p, sc);          // fmin / fmax above swallow a NaN
}

FLM double fl_expm1(double y)
{
    double sc, p;
    flm_exp_parts(y, sc, p);
    return y != y ? y : fma(sc, p, sc - 1.0);
}

// tanh(x) = -em / (2 + em), em = expm1(-2 |x|)
FLM double fl_tanh(double x)
{
    const double a = fmin(fabs(x), 40.0);
    const double em = fl_expm1(-2.0 * a);
    const double t = fl_div(-em, 2.0 + em);
    return x != x ? x : (x < 0.0 ? -t : t);
}
