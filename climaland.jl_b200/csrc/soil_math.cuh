// soil_math.cuh -- branch-free FP64 elementary functions for the soil closures.
//
// The implicit soil path is bound by the FP64 pipe and by instruction issue, not by HBM
// (profiles/): per cell and Newton iteration the van Genuchten closure needs two logs,
// two or three exps and several divisions.  CUDA libm's versions carry slow-path branches
// (denormals, infinities, huge arguments) and immediate-operand moves that the closure
// never needs: its arguments are S in [~1e-8, 1), 1 - S^(1/m) in (0, 1], exponents within
// +-700 and diagonally dominant pivots.  These versions keep <= ~2 ulp accuracy (tested
// against numpy on the GPU, tests/test_cuda_math.py) with straight-line code:
//   rcp   MUFU.RCP64H seed (2^-20) + one second-order Newton step  (3 DFMA)
//   div   rcp + one residual correction                            (5 DFMA/DMUL)
//   log   fdlibm-style: x = 2^k m, s = f/(2+f), odd series in s    (~22 FP64)
//   exp   x = k ln2 + r, degree-11 Chebyshev-economised polynomial (~16 FP64)
//   sqrt  MUFU.RSQ64H seed + Goldschmidt + residual correction     (~8 FP64)
// Coefficients: tools/gen_math_coeffs.py (exp), fdlibm e_log.c constants (log).
#pragma once
#include <math.h>

namespace clb {
namespace fm {

__device__ __forceinline__ double rcp_seed(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

__device__ __forceinline__ double rsqrt_seed(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// 1/x for normal finite x != 0.
__device__ __forceinline__ double rcp(double x)
{
    const double r0 = rcp_seed(x);
    const double e = fma(-x, r0, 1.0);     // |e| <= 2^-20 (measured, tests/test_cuda_math.py)
    const double t = fma(e, e, e);         // e + e^2
    return fma(r0, t, r0);                 // r0 (1 + e + e^2): truncation e^3 = 2^-60, below the rounding
}

// a/b for normal finite b != 0.
__device__ __forceinline__ double div(double a, double b)
{
    const double r0 = rcp_seed(b);
    const double e = fma(-b, r0, 1.0);
    const double t = fma(e, e, e);
    const double r = fma(r0, t, r0);
    const double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// a/b for a divisor whose correctly rounded reciprocal inv_b = RN(1/b) is at hand (a launch constant divided on the
// host): q0 = RN(a inv_b) is within an ulp of a/b, and one exact residual then gives RN(a/b) itself (Markstein's
// theorem; no overflow / underflow in the quotient), i.e. the bits of the IEEE division in 3 dependent FP64 instructions
// instead of the inlined division's MUFU + 8 and its slow-path call.
__device__ __forceinline__ double div_by(double a, double b, double inv_b)
{
    const double q = a * inv_b;
    const double rem = fma(-b, q, a);
    const double r = fma(rem, inv_b, q);
    return (fabs(q) <= 1.7976931348623157e308) ? r : q;  // an infinite (or NaN) quotient has no residual: inf - inf
}

// sqrt(x) for normal finite x > 0; sqrt(0) = 0.
__device__ __forceinline__ double sqrt(double x)
{
    const double y0 = rsqrt_seed(x);
    double g = x * y0;              // ~ sqrt(x)
    double h = 0.5 * y0;            // ~ 1/(2 sqrt(x))
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, x);
    g = fma(d, h, g);
    return (x == 0.0) ? 0.0 : g;
}

// log(x) for normal finite x > 0; log(0) = -inf.  fdlibm e_log.c algorithm.
__device__ __forceinline__ double log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;  // m in [sqrt(1/2), sqrt(2))
    hx |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double m = __hiloint2double(hx, lx);
    const double f = m - 1.0;
    const double s = f * rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    // (double)k without a conversion instruction: 2^52 + 2^31 + k in the mantissa
    const double dk = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;
    const double res = dk * ln2_hi - ((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f);
    return (x > 0.0) ? res : ((x == 0.0) ? -INFINITY : NAN);
}

// exp(x); exact limits: exp(-inf) = 0, underflow to 0 below -708, +inf above 709.7.
__device__ __forceinline__ double exp(double x)
{
    const double L2E = 1.4426950408889634074, ln2_hi = 6.93147180369123816490e-01,
                 ln2_lo = 1.90821492927058770002e-10, MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, L2E, MAGIC);
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -ln2_hi, x);
    r = fma(kd, -ln2_lo, r);
    // Estrin evaluation: half the dependency depth of Horner (the pipe has headroom, latency does not)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(r, 1.0, 1.0);
    const double p23 = fma(r, 0.1666666666666668, 0.5000000000000019);
    const double p45 = fma(r, 0.008333333333319589, 0.04166666666648795);
    const double p67 = fma(r, 0.00019841269890076403, 0.0013888888952352863);
    const double p89 = fma(r, 2.755724088722987e-06, 2.4801485441561313e-05);
    const double pab = fma(r, 2.5110049204818658e-08, 2.763265472252779e-07);
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double p = fma(r8, q2, fma(r4, q1, q0));
    // scale by 2^k through the exponent field (result stays normal for x in [-708, 709])
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    double out = (x < -708.0) ? 0.0 : res;
    out = (x > 709.0) ? INFINITY : out;
    return (x != x) ? x : out;
}

// Variants for arguments known to be in range (the closures: S in (0, 1), exponents within
// +-700 for every admissible parameter set).  log_pos: x normal and > 0, no special cases.
// exp_clamped: no special cases either; the power of two is clamped to the normal range, so
// an out-of-range argument gives a tiny / huge finite number instead of 0 / inf, never garbage.
__device__ __forceinline__ double log_pos(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    hx |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double m = __hiloint2double(hx, lx);
    const double f = m - 1.0;
    const double s = f * rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;
    return dk * ln2_hi - ((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f);
}

__device__ __forceinline__ double exp_clamped(double x)
{
    const double L2E = 1.4426950408889634074, ln2_hi = 6.93147180369123816490e-01,
                 ln2_lo = 1.90821492927058770002e-10, MAGIC = 6755399441055744.0;
    const double t = fma(x, L2E, MAGIC);
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -ln2_hi, x);
    r = fma(kd, -ln2_lo, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(r, 1.0, 1.0);
    const double p23 = fma(r, 0.1666666666666668, 0.5000000000000019);
    const double p45 = fma(r, 0.008333333333319589, 0.04166666666648795);
    const double p67 = fma(r, 0.00019841269890076403, 0.0013888888952352863);
    const double p89 = fma(r, 2.755724088722987e-06, 2.4801485441561313e-05);
    const double pab = fma(r, 2.5110049204818658e-08, 2.763265472252779e-07);
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double p = fma(r8, q2, fma(r4, q1, q0));
    const int kc = min(max(k, -1021), 1022);
    return __hiloint2double(__double2hiint(p) + (kc << 20), __double2loint(p));
}

}  // namespace fm
}  // namespace clb
