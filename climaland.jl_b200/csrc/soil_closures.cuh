// soil_closures.cuh -- device point functions of the implicit soil-column path.
//
// Reference (paths relative to the ClimaLand.jl tree):
//   effective_saturation        src/standalone/Soil/soil_hydrology_parameterizations.jl:45-50
//   pressure_head (vG / BC)     :109-127 / :270-289   (matric_potential :59-63 / :182-186)
//   dψdϑ (vG / BC)              :135-152 / :220-231
//   hydraulic_conductivity      :161-173 / :239-251
//   volumetric_heat_capacity    src/standalone/Soil/soil_heat_parameterizations.jl:157-171
//   temperature_from_ρe_int     :180-192
//   volumetric_internal_energy_liq :221-231
//
// Two arithmetic modes (clb_config.math_mode):
//   LIBM  the reference's expressions written literally with CUDA libm pow().
//   FAST  the same functions with every power of S derived from ONE log(S):
//           A = S^(1/m) = exp(L/m),  B = S^(-1/m) = 1/A,  x = B - 1,
//           (1-A)^m = exp(m*log(1-A)),  x^(1/n) = exp((log(1-A) - L/m)/n),
//           psi = -x^(1/n)/alpha,  dpsi = x^(1/n)*B / (x*S*alpha*m*n*(nu-theta_r)).
//         The FP64 pipe, not HBM, bounds this path, so the number of
//         transcendental evaluations per cell is what the kernel time follows.
// Branch decisions (S < 1, S <= 1, the sqrt(eps) clips) use the same IEEE
// operations on the same inputs as the reference, so they are identical.
#pragma once
#include <math.h>

#include "soil_math.cuh"

namespace clb {

constexpr double kSqrtEps = 1.4901161193847656e-8;  // sqrt(eps(Float64))
constexpr double kEps = 2.220446049250313e-16;      // eps(Float64)

enum { kVanGenuchten = 0, kBrooksCorey = 1 };
enum { kMathFast = 0, kMathLibm = 1 };

// Per-cell hydrology parameters exactly as the reference holds them.
struct HydroCell {
    double nu, theta_r, K_sat, S_s;
    double a;  // vG alpha | BC c
    double b;  // vG n     | BC psi_b
    double m;  // vG m     | unused
};

struct EarthConst {
    double rho_l, rho_i, cp_l, cp_i, T_ref, LH_f0;
};

// Evaluates any subset of {K, psi, dpsi/dtheta} for one cell.  nu_eff is nu for
// Richards and nu - theta_i for EnergyHydrology.  Unused outputs are removed by
// the compiler once inlined.
template <int CLOSURE, int MATH, bool WANT_K, bool WANT_PSI, bool WANT_DPSI>
__device__ __forceinline__ void closure_eval(const HydroCell &p, double theta, double nu_eff,
                                             double &K, double &psi, double &dpsi)
{
    const double theta_safe = fmax(theta, p.theta_r + kSqrtEps);
    const double nu_safe = fmax(nu_eff, p.theta_r + kSqrtEps);
    const double range = nu_safe - p.theta_r;
    const double S = (theta_safe - p.theta_r) / range;

    if (CLOSURE == kVanGenuchten) {
        const double alpha = p.a, n = p.b, m = p.m;
        if (MATH == kMathLibm) {
            if (WANT_K) {
                if (S < 1.0) {
                    double t = 1.0 - pow(1.0 - pow(S, 1.0 / m), m);
                    K = (sqrt(S) * (t * t)) * p.K_sat;
                } else {
                    K = p.K_sat;
                }
            }
            if (WANT_PSI) {
                if (S <= 1.0)
                    psi = -pow((pow(S, -1.0 / m) - 1.0) * pow(alpha, -n), 1.0 / n);
                else
                    psi = (theta_safe - nu_safe) / p.S_s;
            }
            if (WANT_DPSI) {
                if (S < 1.0)
                    dpsi = 1.0 / (alpha * m * n) / range * pow(pow(S, -1.0 / m) - 1.0, 1.0 / n - 1.0) *
                           pow(S, -1.0 / m - 1.0);
                else
                    dpsi = 1.0 / p.S_s;
            }
        } else {
            if (S < 1.0) {
                const double L = log(S);
                const double E = L / m;
                const double A = exp(E);      // S^(1/m)
                const double omA = 1.0 - A;
                const double l1 = log(omA);
                if (WANT_K) {
                    const double t = 1.0 - exp(m * l1);
                    K = (sqrt(S) * (t * t)) * p.K_sat;
                }
                if (WANT_PSI || WANT_DPSI) {
                    const double B = 1.0 / A;           // S^(-1/m)
                    const double x = B - 1.0;
                    const double q = exp((l1 - E) / n); // x^(1/n)
                    if (WANT_PSI) psi = -(q / alpha);
                    if (WANT_DPSI) {
                        const double d = (q * B) / (x * S * (alpha * m * n) * range);
                        dpsi = (x == 0.0) ? INFINITY : d;
                    }
                }
            } else {
                if (WANT_K) K = p.K_sat;
                if (WANT_PSI) psi = (S == 1.0) ? -0.0 : (theta_safe - nu_safe) / p.S_s;
                if (WANT_DPSI) dpsi = 1.0 / p.S_s;
            }
        }
    } else {  // Brooks-Corey
        const double c = p.a, psi_b = p.b;
        if (MATH == kMathLibm) {
            if (WANT_K) K = ((S < 1.0) ? pow(S, 2.0 / c + 3.0) : 1.0) * p.K_sat;
            if (WANT_PSI) {
                if (S <= 1.0)
                    psi = psi_b * pow(S, -1.0 / c);
                else
                    psi = (theta_safe - nu_safe) / p.S_s + psi_b;
            }
            if (WANT_DPSI) {
                if (S < 1.0)
                    dpsi = -psi_b / (c * range) * pow(S, -(1.0 + 1.0 / c));
                else
                    dpsi = 1.0 / p.S_s;
            }
        } else {
            if (S < 1.0) {
                const double L = log(S);
                if (WANT_K) K = exp((2.0 / c + 3.0) * L) * p.K_sat;
                if (WANT_PSI || WANT_DPSI) {
                    const double pw = exp(-L / c);  // S^(-1/c)
                    if (WANT_PSI) psi = psi_b * pw;
                    if (WANT_DPSI) dpsi = -psi_b * pw / (c * range * S);
                }
            } else {
                if (WANT_K) K = p.K_sat;
                if (WANT_PSI) psi = (S == 1.0) ? psi_b : (theta_safe - nu_safe) / p.S_s + psi_b;
                if (WANT_DPSI) dpsi = 1.0 / p.S_s;
            }
        }
    }
}

// Per-cell evaluator of one implicit stage.  nu_eff (nu, or nu - theta_i for
// EnergyHydrology) is constant during the Newton loop, so everything that does not
// depend on the iterate is prepared once: clips, 1/m, 1/n, 1/alpha, the dpsi prefactor.
//   LIBM mode: holds the raw parameters and evaluates the reference's pow expressions.
//   FAST mode: one log(S) shared by all powers, soil_math.cuh functions, no division by a
//              stage constant inside the loop.  The logarithm here is the series version with a
//              RELATIVE error bound (log_pos), not the cheaper table-driven one of the lane and
//              explicit-stage kernels (absolute bound): the reference's hydrostatic test
//              (test/standalone/Soil/soiltest.jl:15-90) asks psi + z to be constant to 2 eps, which
//              the hooks meet and a table logarithm near S = 1 would not (measured).  S itself is an exactly rounded-or-exact
//              quotient (fm::div reproduces exact quotients), so the S < 1 / S <= 1 branches
//              agree with the reference.
template <int CLOSURE, int MATH>
struct CellEval;

template <int CLOSURE>
struct CellEval<CLOSURE, kMathLibm> {
    HydroCell p;
    double nu_eff;
    __device__ __forceinline__ CellEval(const HydroCell &c, double nu_eff_) : p(c), nu_eff(nu_eff_) {}
    template <bool WK, bool WP, bool WD>
    __device__ __forceinline__ void eval(double theta, double &K, double &psi, double &dpsi) const
    {
        closure_eval<CLOSURE, kMathLibm, WK, WP, WD>(p, theta, nu_eff, K, psi, dpsi);
    }
};

template <>
struct CellEval<kVanGenuchten, kMathFast> {
    double theta_r, theta_lo, nu_safe, range, K_sat, inv_Ss, m, inv_m, inv_n, inv_alpha, c_dpsi;
    __device__ __forceinline__ CellEval(const HydroCell &c, double nu_eff)
    {
        theta_r = c.theta_r;
        theta_lo = c.theta_r + kSqrtEps;
        nu_safe = fmax(nu_eff, theta_lo);
        range = nu_safe - c.theta_r;
        K_sat = c.K_sat;
        inv_Ss = fm::rcp(c.S_s);
        m = c.m;
        inv_m = fm::rcp(c.m);
        inv_n = fm::rcp(c.b);
        inv_alpha = fm::rcp(c.a);
        c_dpsi = fm::rcp((c.a * c.m * c.b) * range);
    }
    template <bool WK, bool WP, bool WD>
    __device__ __forceinline__ void eval(double theta, double &K, double &psi, double &dpsi) const
    {
        const double theta_safe = fmax(theta, theta_lo);
        const double S = fm::div(theta_safe - theta_r, range);
        if (S < 1.0) {
            const double L = fm::log_pos(S);      // S in [~1e-8, 1)
            const double E = L * inv_m;
            const double A = fm::exp_clamped(E);  // S^(1/m) in (0, 1]
            // 1 - A is 0 only when S^(1/m) rounds to 1; a floor keeps log finite (the limits
            // K -> K_sat, psi -> 0 are reached either way, dpsi is selected below)
            const double l1 = fm::log_pos(fmax(1.0 - A, 1e-300));
            if (WK) {
                const double t = 1.0 - fm::exp_clamped(m * l1);
                K = (fm::sqrt(S) * (t * t)) * K_sat;
            }
            if (WP || WD) {
                const double B = fm::rcp(A);  // S^(-1/m)
                const double x = B - 1.0;
                const double q = fm::exp_clamped((l1 - E) * inv_n);  // x^(1/n)
                if (WP) psi = -(q * inv_alpha);
                if (WD) {
                    const double d = ((q * B) * c_dpsi) * fm::rcp(x * S);
                    dpsi = (x == 0.0) ? INFINITY : d;
                }
            }
        } else {
            if (WK) K = K_sat;
            if (WP) psi = (S == 1.0) ? -0.0 : (theta_safe - nu_safe) * inv_Ss;
            if (WD) dpsi = inv_Ss;
        }
    }
};

template <>
struct CellEval<kBrooksCorey, kMathFast> {
    double theta_r, theta_lo, nu_safe, range, K_sat, inv_Ss, psi_b, neg_inv_c, k_exp, c_dpsi;
    __device__ __forceinline__ CellEval(const HydroCell &c, double nu_eff)
    {
        theta_r = c.theta_r;
        theta_lo = c.theta_r + kSqrtEps;
        nu_safe = fmax(nu_eff, theta_lo);
        range = nu_safe - c.theta_r;
        K_sat = c.K_sat;
        inv_Ss = fm::rcp(c.S_s);
        psi_b = c.b;
        neg_inv_c = -fm::rcp(c.a);
        k_exp = fma(-2.0, neg_inv_c, 3.0);
        c_dpsi = -fm::div(c.b, c.a * range);
    }
    template <bool WK, bool WP, bool WD>
    __device__ __forceinline__ void eval(double theta, double &K, double &psi, double &dpsi) const
    {
        const double theta_safe = fmax(theta, theta_lo);
        const double S = fm::div(theta_safe - theta_r, range);
        if (S < 1.0) {
            const double L = fm::log_pos(S);
            if (WK) K = fm::exp_clamped(k_exp * L) * K_sat;
            if (WP || WD) {
                const double pw = fm::exp_clamped(L * neg_inv_c);  // S^(-1/c)
                if (WP) psi = psi_b * pw;
                if (WD) dpsi = (c_dpsi * pw) * fm::rcp(S);
            }
        } else {
            if (WK) K = K_sat;
            if (WP) psi = (S == 1.0) ? psi_b : (theta_safe - nu_safe) * inv_Ss + psi_b;
            if (WD) dpsi = inv_Ss;
        }
    }
};

template <int CLOSURE, int MATH>
__device__ __forceinline__ double pressure_head(const HydroCell &p, double theta, double nu_eff)
{
    double K, psi, d;
    CellEval<CLOSURE, MATH>(p, nu_eff).template eval<false, true, false>(theta, K, psi, d);
    return psi;
}

template <int CLOSURE, int MATH>
__device__ __forceinline__ double dpsidtheta(const HydroCell &p, double theta, double nu_eff)
{
    double K, psi, d;
    CellEval<CLOSURE, MATH>(p, nu_eff).template eval<false, false, true>(theta, K, psi, d);
    return d;
}

// soil_heat_parameterizations.jl:157-171
__device__ __forceinline__ double volumetric_heat_capacity(double theta_l, double theta_i, double rho_c_ds,
                                                           const EarthConst &e)
{
    return rho_c_ds + theta_l * (e.cp_l * e.rho_l) + theta_i * (e.cp_i * e.rho_i);
}

// soil_heat_parameterizations.jl:180-192
__device__ __forceinline__ double temperature_from_rho_e_int(double rho_e_int, double theta_i, double rho_c_s,
                                                             const EarthConst &e)
{
    return e.T_ref + (rho_e_int + theta_i * e.rho_i * e.LH_f0) / rho_c_s;
}

// soil_heat_parameterizations.jl:221-231
__device__ __forceinline__ double volumetric_internal_energy_liq(double T, const EarthConst &e)
{
    return (e.cp_l * e.rho_l) * (T - e.T_ref);
}

// update_implicit_aux of EnergyHydrology (energy_hydrology.jl:427-445): T from
// (theta_l clipped to the pore space left by ice, rho_e_int, theta_i).
__device__ __forceinline__ double eh_temperature(double theta_l, double rho_e_int, double theta_i, double nu,
                                                 double rho_c_ds, const EarthConst &e)
{
    const double tl = fmin(nu - theta_i, theta_l);
    const double rho_c_s = volumetric_heat_capacity(tl, theta_i, rho_c_ds, e);
    return temperature_from_rho_e_int(rho_e_int, theta_i, rho_c_s, e);
}

// The same with the branch-free division of soil_math.cuh (FAST mode of the fused kernels).
template <int MATH>
__device__ __forceinline__ double eh_temperature_m(double theta_l, double rho_e_int, double theta_i, double nu,
                                                   double rho_c_ds, const EarthConst &e)
{
    if (MATH == kMathLibm) return eh_temperature(theta_l, rho_e_int, theta_i, nu, rho_c_ds, e);
    const double tl = fmin(nu - theta_i, theta_l);
    const double rho_c_s = volumetric_heat_capacity(tl, theta_i, rho_c_ds, e);
    return e.T_ref + fm::div(rho_e_int + theta_i * e.rho_i * e.LH_f0, rho_c_s);
}

}  // namespace clb
