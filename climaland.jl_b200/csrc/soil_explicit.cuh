// soil_explicit.cuh -- the explicit-stage soil kernels of EnergyHydrology: the step
// immediately before the implicit solve.  They PRODUCE the lagged inputs of the implicit stage (p.soil.K,
// kappa, theta_l) in the library's mirrors, so with them on the device the host uploads only the state.
//
// Reference (paths relative to the ClimaLand.jl tree):
//   update_aux!(p, Y, t)                 src/standalone/Soil/energy_hydrology.jl:722-814
//   source!(dY, ::PhaseChange, Y, p, m)  energy_hydrology.jl:846-906
//   volumetric_liquid_fraction           src/standalone/Soil/soil_hydrology_parameterizations.jl:22-31
//   impedance_factor, viscosity_factor   soil_hydrology_parameterizations.jl:302-305, 320-324
//   matric_potential / inverse           :59-63, 72-77 (van Genuchten), :182-186, 195-200 (Brooks-Corey)
//   soil_Tf_depressed, thermal_time, phase_change_source   src/standalone/Soil/soil_heat_parameterizations.jl:35-122
//   kappa_sat, thermal_conductivity, relative_saturation, kersten_number   soil_heat_parameterizations.jl:243-323
//   total_liq_water_vol_per_area!, total_energy_per_area!   energy_hydrology.jl:1282-1327
//
// One thread per cell (grid: column blocks x levels), every field read once, coalesced in either mirror
// layout's fast direction when the mirrors are column-fastest; the two column integrals are a second,
// thread-per-column kernel (3 reads per cell).  Pointwise FP64 work: ~17 powers per cell.
//   LIBM mode: the reference's pow / exp / log expressions literally with CUDA libm.
//   FAST mode: x^y = exp(y log x) with the branch-free functions of soil_math.cuh (<= ~4e-15 relative for
//              the exponents that occur); zero bases are handled explicitly.
#pragma once
#include "soil_device.cuh"

namespace clb {

// scalars of EnergyHydrologyParameters (energy_hydrology.jl:150-160) + LandParameters T_freeze, grav
struct ExplicitConst {
    double Omega, gamma, gammaT_ref, alpha, beta, T_freeze, grav;
};

// per-cell fields only the explicit stage touches
struct ExplicitView {
    const double *kappa_dry, *kappa_sat_unfrozen, *kappa_sat_frozen, *nu_ss_om, *nu_ss_quartz, *nu_ss_gravel;
    double *p_theta_l, *p_kappa, *p_K, *p_T, *p_psi, *p_Tf;  // p.soil.{theta_l, kappa, K, T, psi, Tf_depressed}
    double *total_water, *total_energy;                     // per column
    double *dYe_theta_l, *dYe_theta_i;                      // explicit tendency the PhaseChange source adds into
    ExplicitConst k;
};

// x^y for x >= 0 (y finite): the reference's Float64 ^ Float64
template <int MATH>
__device__ __forceinline__ double pw(double x, double y)
{
    if (MATH == kMathLibm) return pow(x, y);
    if (x == 0.0) return (y > 0.0) ? 0.0 : ((y == 0.0) ? 1.0 : INFINITY);
    return fm::exp(y * fm::log(x));
}
template <int MATH>
__device__ __forceinline__ double ex(double x) { return (MATH == kMathLibm) ? exp(x) : fm::exp(x); }
template <int MATH>
__device__ __forceinline__ double lg(double x) { return (MATH == kMathLibm) ? log(x) : fm::log(x); }

// soil_hydrology_parameterizations.jl:59-63 / :182-186
template <int CLOSURE, int MATH>
__device__ __forceinline__ double matric_potential(const HydroCell &p, double S)
{
    if (CLOSURE == kVanGenuchten) return -pw<MATH>((pw<MATH>(S, -1.0 / p.m) - 1.0) * pw<MATH>(p.a, -p.b), 1.0 / p.b);
    return p.b * pw<MATH>(S, -1.0 / p.a);
}

// soil_hydrology_parameterizations.jl:72-77 / :195-200 (psi > 0 is an error upstream: NaN here)
template <int CLOSURE, int MATH>
__device__ __forceinline__ double inverse_matric_potential(const HydroCell &p, double psi)
{
    if (psi > 0.0) return NAN;
    if (CLOSURE == kVanGenuchten) return pw<MATH>(1.0 + pw<MATH>(p.a * fabs(psi), p.b), -p.m);
    return pw<MATH>(psi / p.b, -p.a);
}

__device__ __forceinline__ double effective_saturation(double nu_eff, double theta, double theta_r)
{
    const double theta_safe = fmax(theta, theta_r + kSqrtEps);
    const double nu_safe = fmax(nu_eff, theta_r + kSqrtEps);
    return (theta_safe - theta_r) / (nu_safe - theta_r);
}

// soil_heat_parameterizations.jl:35-52; rho_first / rho_second are the positional (_rho_ice, _rho_liq):
// update_aux! passes (rho_l, rho_i), PhaseChange passes (rho_i, rho_l) -- both restated as evaluated there.
template <int CLOSURE, int MATH>
__device__ __forceinline__ double soil_Tf_depressed(const HydroCell &p, double theta_l, double theta_i, double rho_first,
                                                    double rho_second, const ExplicitConst &k, double LH_f0,
                                                    double &psi_w0)
{
    const double theta_tot = fmin(rho_first / rho_second * theta_i + theta_l, p.nu);
    psi_w0 = matric_potential<CLOSURE, MATH>(p, effective_saturation(p.nu, theta_tot, p.theta_r));
    return fmax(k.T_freeze * ex<MATH>(k.grav * psi_w0 / LH_f0), 1.0);
}

// soil_heat_parameterizations.jl:243-254
template <int MATH>
__device__ __forceinline__ double kappa_sat(double theta_l, double theta_i, double ku, double kf)
{
    const double theta_w = theta_l + theta_i;
    if (theta_w < kEps) return (ku + kf) / 2.0;
    return pw<MATH>(ku, theta_l / theta_w) * pw<MATH>(kf, theta_i / theta_w);
}

// soil_heat_parameterizations.jl:301-323
template <int MATH>
__device__ __forceinline__ double kersten_number(double theta_i, double S_r, double alpha, double beta, double om,
                                                 double quartz, double gravel)
{
    if (theta_i < kEps)
        return pw<MATH>(S_r, (1.0 + om - alpha * quartz - gravel) / 2.0) *
               pw<MATH>(pw<MATH>(1.0 + ex<MATH>(-beta * S_r), -3.0) - pw<MATH>((1.0 - S_r) / 2.0, 3.0), 1.0 - om);
    return pw<MATH>(S_r, 1.0 + om);
}

// heaviside(x): src/shared_utilities/utils.jl:83-99 (1 if x > eps, else 0)
__device__ __forceinline__ double heaviside(double x) { return (x > kEps) ? 1.0 : 0.0; }

// AUX: update_aux! for the cell; PHASE: the PhaseChange source of the cell.  With both, the source uses the
// freshly computed theta_l, kappa, T (what the reference reads back from p), so no field is read twice.
template <int CLOSURE, int MATH, bool AUX, bool PHASE>
__global__ void __launch_bounds__(128) k_explicit_cells(const DevView P, const ExplicitView X)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c >= P.ncol) return;
    const int64_t q = P.at(i, c);
    const EarthConst &E = P.earth;
    const HydroCell cell = load_cell(P, q);
    const double th = P.Y_theta_l[q], thi = P.Y_theta_i[q];
    const double rcds = __ldg(P.rho_c_ds + q);
    double theta_l, kappa, T;
    if (AUX) {
        // theta_l = volumetric_liquid_fraction(theta_l, nu - theta_i, theta_r)
        {
            const double theta_safe = fmax(th, cell.theta_r + kSqrtEps);
            const double nu_safe = fmax(cell.nu - thi, cell.theta_r + kSqrtEps);
            theta_l = (theta_safe < nu_safe) ? theta_safe : nu_safe;
        }
        const double S_r = (theta_l + thi) / cell.nu;
        const double K_e = kersten_number<MATH>(thi, S_r, X.k.alpha, X.k.beta, __ldg(X.nu_ss_om + q),
                                                __ldg(X.nu_ss_quartz + q), __ldg(X.nu_ss_gravel + q));
        const double ks = kappa_sat<MATH>(theta_l, thi, __ldg(X.kappa_sat_unfrozen + q), __ldg(X.kappa_sat_frozen + q));
        kappa = K_e * ks + (1.0 - K_e) * __ldg(X.kappa_dry + q);
        T = temperature_from_rho_e_int(P.Y_rho_e[q], thi, volumetric_heat_capacity(theta_l, thi, rcds, E), E);
        // K = impedance * viscosity * hydraulic_conductivity(effective_saturation(nu, theta_l(Y), theta_r))
        double Kh, psi, d;
        CellEval<CLOSURE, MATH>(cell, cell.nu).template eval<true, false, false>(th, Kh, psi, d);
        const double imp = pw<MATH>(10.0, -X.k.Omega * (thi / (theta_l + thi - cell.theta_r)));
        const double visc = ex<MATH>(X.k.gamma * (T - X.k.gammaT_ref));
        CellEval<CLOSURE, MATH>(cell, cell.nu - thi).template eval<false, true, false>(th, d, psi, d);
        double psi_w0;
        X.p_theta_l[q] = theta_l;
        X.p_kappa[q] = kappa;
        X.p_T[q] = T;
        X.p_K[q] = imp * visc * Kh;
        X.p_psi[q] = psi;
        X.p_Tf[q] = soil_Tf_depressed<CLOSURE, MATH>(cell, theta_l, thi, E.rho_l, E.rho_i, X.k, E.LH_f0, psi_w0);
    } else {
        theta_l = X.p_theta_l[q];
        kappa = X.p_kappa[q];
        T = X.p_T[q];
    }
    if (PHASE) {
        const double dz = P.dz_c[i];
        const double tau = 3.0 * volumetric_heat_capacity(theta_l, thi, rcds, E) * (dz * dz) / kappa;
        double psi_w0;
        const double Tf = soil_Tf_depressed<CLOSURE, MATH>(cell, theta_l, thi, E.rho_i, E.rho_l, X.k, E.LH_f0, psi_w0);
        const double psi_T = E.LH_f0 / X.k.grav * lg<MATH>(T / Tf) * heaviside(Tf - T);
        const double theta_star =
            inverse_matric_potential<CLOSURE, MATH>(cell, psi_w0 + psi_T) * (cell.nu - cell.theta_r) + cell.theta_r;
        const double s = (theta_l - theta_star) / tau;
        X.dYe_theta_l[q] += -s;
        X.dYe_theta_i[q] += (E.rho_l / E.rho_i) * s;
    }
}

// total_liq_water_vol_per_area! and total_energy_per_area! (ClimaCore column_integral_definite!)
__global__ void __launch_bounds__(128) k_explicit_totals(const DevView P, const ExplicitView X)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const EarthConst &E = P.earth;
    double tw = 0.0, te = 0.0;
    for (int i = 0; i < P.N; ++i) {
        const int64_t q = P.at(i, c);
        const double dz = P.dz_c[i];
        tw += (P.Y_theta_l[q] + P.Y_theta_i[q] * E.rho_i / E.rho_l) * dz;
        te += P.Y_rho_e[q] * dz;
    }
    X.total_water[c] = tw;
    X.total_energy[c] = te;
}

}  // namespace clb
