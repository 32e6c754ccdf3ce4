// soil_mathv.cuh -- the branch-free FP64 functions of soil_math.cuh and the soil closures on W
// independent arguments at once, written statement by statement across the W values.
//
// Why: a DFMA has 8 cycles of latency and the FP64 pipe of an SM sub-partition accepts one
// warp-instruction every 2 cycles (tools/ubench/fp64_ilp.cu), so it needs >= 4 independent FP64
// instructions in flight.  The lane-pair kernel runs 2 warps per sub-partition (its registers hold
// eight cells per lane), so the parallelism has to come from inside the thread: every line below
// is W independent instructions in program order, which is the order ptxas keeps.
// Same algorithms, coefficients and rounding sequence per value as soil_math.cuh / CellEval.
#pragma once
#include "soil_closures.cuh"
#include "soil_math_tables.cuh"

namespace clb {
namespace fmv {

#define CLB_V _Pragma("unroll") for (int j = 0; j < W; ++j)

__device__ __forceinline__ double max_nn(double a, double b) { return (a > b) ? a : b; }  // b when a is NaN (as fmax)
__device__ __forceinline__ double min_nn(double a, double b) { return (a < b) ? a : b; }

template <int W>
__device__ __forceinline__ void rcp(const double (&x)[W], double (&r)[W])
{
    double r0[W], e[W];
    CLB_V r0[j] = fm::rcp_seed(x[j]);
    CLB_V e[j] = fma(-x[j], r0[j], 1.0);
    CLB_V e[j] = fma(e[j], e[j], e[j]);
    CLB_V r[j] = fma(r0[j], e[j], r0[j]);
}

template <int W>
__device__ __forceinline__ void div(const double (&a)[W], const double (&b)[W], double (&q)[W])
{
    double r[W], rem[W];
    rcp<W>(b, r);
    CLB_V q[j] = a[j] * r[j];
    CLB_V rem[j] = fma(-b[j], q[j], a[j]);
    CLB_V q[j] = fma(rem[j], r[j], q[j]);
}

template <int W>
__device__ __forceinline__ void sqrt(const double (&x)[W], double (&out)[W])
{
    double g[W], h[W], e[W];
    CLB_V {
        const double y0 = fm::rsqrt_seed(x[j]);
        g[j] = x[j] * y0;
        h[j] = 0.5 * y0;
    }
    CLB_V e[j] = fma(-g[j], h[j], 0.5);
    CLB_V { g[j] = fma(g[j], e[j], g[j]); h[j] = fma(h[j], e[j], h[j]); }
    CLB_V e[j] = fma(-g[j], h[j], 0.5);
    CLB_V { g[j] = fma(g[j], e[j], g[j]); h[j] = fma(h[j], e[j], h[j]); }
    CLB_V e[j] = fma(-g[j], g[j], x[j]);
    CLB_V g[j] = fma(e[j], h[j], g[j]);
    CLB_V out[j] = (x[j] == 0.0) ? 0.0 : g[j];
}

// log(x), x normal and > 0 (fm::log_pos)
template <int W>
__device__ __forceinline__ void log_pos(const double (&x)[W], double (&out)[W])
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    double f[W], dk[W], d2[W], r[W], s[W], z[W], w[W], t1[W], t2[W], hfsq[W];
    CLB_V {
        int hx = __double2hiint(x[j]);
        const int lx = __double2loint(x[j]);
        int k = (hx >> 20) - 1023;
        hx &= 0x000fffff;
        const int i = (hx + 0x95f64) & 0x100000;
        hx |= (i ^ 0x3ff00000);
        k += (i >> 20);
        f[j] = __hiloint2double(hx, lx) - 1.0;
        dk[j] = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;
    }
    CLB_V d2[j] = 2.0 + f[j];
    rcp<W>(d2, r);
    CLB_V s[j] = f[j] * r[j];
    CLB_V hfsq[j] = 0.5 * f[j] * f[j];
    CLB_V z[j] = s[j] * s[j];
    CLB_V w[j] = z[j] * z[j];
    CLB_V t1[j] = fma(w[j], Lg6, Lg4);
    CLB_V t2[j] = fma(w[j], Lg7, Lg5);
    CLB_V t1[j] = fma(w[j], t1[j], Lg2);
    CLB_V t2[j] = fma(w[j], t2[j], Lg3);
    CLB_V t1[j] = w[j] * t1[j];
    CLB_V t2[j] = fma(w[j], t2[j], Lg1);
    CLB_V t2[j] = z[j] * t2[j];
    CLB_V t1[j] = t2[j] + t1[j];                                    // R
    CLB_V t1[j] = fma(s[j], hfsq[j] + t1[j], dk[j] * ln2_lo);
    CLB_V out[j] = dk[j] * ln2_hi - ((hfsq[j] - t1[j]) - f[j]);
}

// exp(x), no special cases, power of two clamped to the normal range (fm::exp_clamped)
template <int W>
__device__ __forceinline__ void exp_clamped(const double (&x)[W], double (&out)[W])
{
    const double L2E = 1.4426950408889634074, ln2_hi = 6.93147180369123816490e-01,
                 ln2_lo = 1.90821492927058770002e-10, MAGIC = 6755399441055744.0;
    double t[W], r[W], r2[W], r4[W], r8[W], q0[W], q1[W], q2[W];
    int k[W];
    CLB_V t[j] = fma(x[j], L2E, MAGIC);
    CLB_V {
        k[j] = __double2loint(t[j]);
        t[j] = t[j] - MAGIC;
    }
    CLB_V r[j] = fma(t[j], -ln2_hi, x[j]);
    CLB_V r[j] = fma(t[j], -ln2_lo, r[j]);
    CLB_V r2[j] = r[j] * r[j];
    CLB_V q0[j] = fma(r[j], 0.1666666666666668, 0.5000000000000019);
    CLB_V q1[j] = fma(r[j], 0.00019841269890076403, 0.0013888888952352863);
    CLB_V q2[j] = fma(r[j], 2.5110049204818658e-08, 2.763265472252779e-07);
    CLB_V r4[j] = r2[j] * r2[j];
    CLB_V q0[j] = fma(r2[j], q0[j], r[j] + 1.0);
    CLB_V q1[j] = fma(r2[j], q1[j], fma(r[j], 0.008333333333319589, 0.04166666666648795));
    CLB_V q2[j] = fma(r2[j], q2[j], fma(r[j], 2.755724088722987e-06, 2.4801485441561313e-05));
    CLB_V r8[j] = r4[j] * r4[j];
    CLB_V q0[j] = fma(r4[j], q1[j], q0[j]);
    CLB_V q0[j] = fma(r8[j], q2[j], q0[j]);
    CLB_V {
        const int kc = min(max(k[j], -1021), 1022);
        out[j] = __hiloint2double(__double2hiint(q0[j]) + (kc << 20), __double2loint(q0[j]));
    }
}

// ---- table-driven log / exp (tables: soil_math_tables.cuh, copied to shared memory per block) -----
// The closures use a logarithm only as (a factor of) the argument of an exp, so what they need is
// an ABSOLUTE error of ~2^-56 max(1, |log x|), not a relative one near x = 1; that is what one table
// look-up and a degree-6 series give at less than half the FP64 instructions of the series-only
// versions (log 11 against 26, exp 10 against 17; tests/test_cuda_math.py states the bounds).
struct MathTab {
    const double2 *logt;  // [128] { 1/c, ln c }
    const double *expt;   // [64] 2^(j/64)
};
constexpr int kMathTabBytes = mtab::kLogN * 16 + mtab::kExpN * 8;

// the first kLogN threads of a block fill the tables; the caller synchronises the block
__device__ __forceinline__ MathTab math_tab_fill(unsigned char *smem, int tid)
{
    double2 *lt = reinterpret_cast<double2 *>(smem);
    double *et = reinterpret_cast<double *>(smem + mtab::kLogN * 16);
    if (tid < mtab::kLogN) lt[tid] = mtab::g_log_tab[tid];
    if (tid < mtab::kExpN) et[tid] = mtab::g_exp_tab[tid];
    return MathTab{lt, et};
}

// log(x), x normal and > 0: x = 2^k z, z in [0.6875, 1.375); ln z = ln c + log1p(r), r = z/c - 1 by one fma
template <int W>
__device__ __forceinline__ void log_tab(const MathTab &T, const double (&x)[W], double (&out)[W])
{
    double z[W], dk[W], r[W], w[W], r2[W], p01[W], p23[W];
    double2 e[W];
    CLB_V {
        const int hx = __double2hiint(x[j]);
        const int tmp = hx - mtab::kLogOffHi;
        const int k = tmp >> 20;
                e[j] = T.logt[(tmp >> 13) & (mtab::kLogN - 1)];
        z[j] = __hiloint2double(hx - (tmp & 0xfff00000), __double2loint(x[j]));
        // the exponent as a double: one conversion (I2F.F64) instead of the magic-number LOP3 + DADD of the series
        // version above -- exact either way; the lane kernels pay 1 cycle per non-FP64 and 2 per FP64 instruction
        // (DESIGN.md section 10): 48.8 -> 47.9 us (EnergyHydrology), 31.0 -> 30.4 us (Richards) per ~1 degree stage
        dk[j] = __int2double_rn(k);
    }
    CLB_V r[j] = fma(z[j], e[j].x, -1.0);
    CLB_V w[j] = fma(dk[j], mtab::kLn2, e[j].y);
    CLB_V r2[j] = r[j] * r[j];
    CLB_V p01[j] = fma(r[j], 0.33333333333333333, -0.5);
    CLB_V p23[j] = fma(r[j], 0.2, -0.25);
    CLB_V p23[j] = fma(r2[j], -0.16666666666666667, p23[j]);
    CLB_V p01[j] = fma(r2[j], p23[j], p01[j]);
    CLB_V r[j] = fma(r2[j], p01[j], r[j]);
    CLB_V out[j] = w[j] + r[j];
}

// exp(x), no special cases, power of two clamped to the normal range: x = (64 e + i) ln2/64 + r
template <int W>
__device__ __forceinline__ void exp_tab(const MathTab &T, const double (&x)[W], double (&out)[W])
{
    const double MAGIC = 6755399441055744.0;
    double t[W], r[W], r2[W], q23[W], q45[W], tb[W];
    int k[W];
    CLB_V t[j] = fma(x[j], mtab::kExpScale, MAGIC);
    CLB_V {
        k[j] = __double2loint(t[j]);
        t[j] = t[j] - MAGIC;
                tb[j] = T.expt[k[j] & (mtab::kExpN - 1)];
    }
    CLB_V r[j] = fma(t[j], -mtab::kLn2NHi, x[j]);
    CLB_V r[j] = fma(t[j], -mtab::kLn2NLo, r[j]);
    CLB_V r2[j] = r[j] * r[j];
    CLB_V q23[j] = fma(r[j], mtab::kExpC3, mtab::kExpC2);
    CLB_V q45[j] = fma(r[j], mtab::kExpC5, mtab::kExpC4);
    CLB_V q23[j] = fma(r2[j], q45[j], q23[j]);
    CLB_V r[j] = fma(r2[j], q23[j], r[j]);
    CLB_V r[j] = fma(tb[j], r[j], tb[j]);
    CLB_V {
        const int kc = min(max(k[j] >> 6, -1021), 1022);
        out[j] = __hiloint2double(__double2hiint(r[j]) + (kc << 20), __double2loint(r[j]));
    }
}

template <bool TAB, int W>
__device__ __forceinline__ void log_any(const MathTab &T, const double (&x)[W], double (&out)[W])
{
    if constexpr (TAB) log_tab<W>(T, x, out);
    else log_pos<W>(x, out);
}
template <bool TAB, int W>
__device__ __forceinline__ void exp_any(const MathTab &T, const double (&x)[W], double (&out)[W])
{
    if constexpr (TAB) exp_tab<W>(T, x, out);
    else exp_clamped<W>(x, out);
}

// K (optional), psi and dpsi/dtheta of W cells (soil_hydrology_parameterizations.jl:45-50, 109-173,
// 220-289).  Branch-free: the unsaturated formulas run on every value with finite garbage on saturated
// ones; the S < 1 / S == 1 decisions are selects on the same IEEE comparisons as the reference.
// Constants: ClosureConst of soil_pair.cuh (van Genuchten ca = 1/m, ca2 = m, cb = 1/n, cc = 1/alpha,
// cd = 1/(alpha m n range); Brooks-Corey ca = -1/c, ca2 = 2/c + 3, cb = psi_b, cc = -psi_b/(c range)).
template <int CLOSURE, bool WK, int W, bool TAB = false>
__device__ __forceinline__ void closure(const MathTab &T, const double (&theta)[W], const double (&theta_r)[W], const double (&nu_eff)[W],
                                        const double (&inv_range)[W], const double (&ca)[W], const double (&ca2)[W], const double (&cb)[W],
                                        const double (&cc)[W], const double (&cd)[W], const double (&inv_Ss)[W],
                                        const double (&K_sat)[W], double (&K)[W], double (&psi)[W], double (&dps)[W])
{
    // effective saturation: the quotient num / range of the reference is taken as num * (1/range) with the
    // stage-constant reciprocal (<= 1.5 ulp instead of 0.5 ulp); its comparisons with 1 are decided exactly
    // as the reference's: num and range are doubles, so the correctly rounded quotient is < 1 iff num < range
    // and == 1 iff num == range, and th_safe - theta_r vs nu_safe - theta_r compare as th_safe vs nu_safe
    // except when both differences round to the same double, which the explicit differences below keep.
    double lo[W], nu_safe[W], range[W], th_safe[W], num[W], S[W], L[W];
    bool unsat[W];
    CLB_V lo[j] = theta_r[j] + kSqrtEps;
    CLB_V nu_safe[j] = max_nn(nu_eff[j], lo[j]);
    CLB_V th_safe[j] = max_nn(theta[j], lo[j]);
    CLB_V range[j] = nu_safe[j] - theta_r[j];
    CLB_V num[j] = th_safe[j] - theta_r[j];
    CLB_V S[j] = num[j] * inv_range[j];
#ifdef CLB_EXACT_S
    {
        // one residual correction: the correctly rounded quotient num / range (as the reference's division) in all but
        // ~1e-16 of the cases instead of <= 1.5 ulp
        double rem[W];
        CLB_V rem[j] = fma(-range[j], S[j], num[j]);
        CLB_V S[j] = fma(rem[j], inv_range[j], S[j]);
    }
#endif
    CLB_V unsat[j] = num[j] < range[j];
    log_any<TAB, W>(T, S, L);
    if (CLOSURE == kVanGenuchten) {
        double Ee[W], A[W], omA[W], arg[W], l1[W], qn[W], den[W], rd[W];
        CLB_V Ee[j] = L[j] * ca[j];
        exp_any<TAB, W>(T, Ee, A);  // S^(1/m)
        CLB_V omA[j] = 1.0 - A[j];
        // 1 - A is 0 only when S^(1/m) rounds to 1; a floor keeps log finite (dpsi is selected below)
        // 0 <= omA <= 1: adding 1e-300 is the floor max(omA, 1e-300) without a compare and two selects
        CLB_V arg[j] = omA[j] + 1e-300;
        log_any<TAB, W>(T, arg, l1);
        if (WK) {
            double em[W], t[W], sq[W];
            CLB_V em[j] = ca2[j] * l1[j];
            exp_any<TAB, W>(T, em, t);
            sqrt<W>(S, sq);
            CLB_V t[j] = 1.0 - t[j];
            CLB_V K[j] = unsat[j] ? (sq[j] * (t[j] * t[j])) * K_sat[j] : K_sat[j];
        }
        // (S^(-1/m) - 1)^(1/n) = ((1 - A)/A)^(1/n);  dpsi = that / ((1 - A) S alpha m n range)
        CLB_V arg[j] = (l1[j] - Ee[j]) * cb[j];
        exp_any<TAB, W>(T, arg, qn);
        CLB_V den[j] = omA[j] * S[j];
        rcp<W>(den, rd);
        CLB_V {
            // S == 1: the reference's matric potential is -0.0 and this expression +0.0; psi only enters as
            // psi + z here, where the sign of a zero is lost, so the S == 1 select of CellEval is not needed
            const double psi_s = (th_safe[j] - nu_safe[j]) * inv_Ss[j];
            psi[j] = unsat[j] ? -(qn[j] * cc[j]) : psi_s;
            double d = (qn[j] * cd[j]) * rd[j];
            d = (omA[j] <= 0.0) ? INFINITY : d;
            dps[j] = unsat[j] ? d : inv_Ss[j];
        }
    } else {
        double arg[W], pw[W], rS[W];
        if (WK) {
            double ek[W], t[W];
            CLB_V ek[j] = ca2[j] * L[j];
            exp_any<TAB, W>(T, ek, t);
            CLB_V K[j] = unsat[j] ? t[j] * K_sat[j] : K_sat[j];
        }
        CLB_V arg[j] = L[j] * ca[j];
        exp_any<TAB, W>(T, arg, pw);  // S^(-1/c)
        rcp<W>(S, rS);
        CLB_V {
            const double psi_s = (th_safe[j] - nu_safe[j]) * inv_Ss[j] + cb[j];  // S == 1: 0 * inv_Ss + psi_b = psi_b
            psi[j] = unsat[j] ? cb[j] * pw[j] : psi_s;
            dps[j] = unsat[j] ? (cc[j] * pw[j]) * rS[j] : inv_Ss[j];
        }
    }
}

#undef CLB_V

}  // namespace fmv
}  // namespace clb
