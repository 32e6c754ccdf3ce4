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
//              the exponents that occur); zero bases are handled explicitly; integer powers are products,
//              10^x = exp(x ln 10), products of powers share one exp, and a cell without ice evaluates the
//              closure and the depressed freezing point once instead of twice (theta_i = 0 makes the two
//              saturations / the two ice-to-liquid ratios coincide).
#pragma once
#include "soil_device.cuh"
#include "soil_mathv.cuh"

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
    int assign_source = 0;                                  // 1: the source is stored, not added (clb_soil_step_host)
};

// Table-driven log / exp of soil_mathv.cuh (11 / 10 FP64 instructions against 26 / 17 of the series-only
// versions), reading the 2.5 KB of tables through L1 from global memory.  tlog: x normal and > 0, absolute
// error ~2^-56 max(1, |log x|) -- what a power needs, not a relative bound near 1; texp: no special cases, the
// power of two is clamped to the normal range.
__device__ __forceinline__ double tlog(double x)
{
    const fmv::MathTab MT{mtab::g_log_tab, mtab::g_exp_tab};
    const double xv[1] = {x};
    double r[1];
    fmv::log_tab<1>(MT, xv, r);
    return r[0];
}
__device__ __forceinline__ double texp(double x)
{
    const fmv::MathTab MT{mtab::g_log_tab, mtab::g_exp_tab};
    const double xv[1] = {x};
    double r[1];
    fmv::exp_tab<1>(MT, xv, r);
    return r[0];
}

// x^y for x >= 0 (y finite): the reference's Float64 ^ Float64
template <int MATH>
__device__ __forceinline__ double pw(double x, double y)
{
    if (MATH == kMathLibm) return pow(x, y);
    if (x == 0.0) return (y > 0.0) ? 0.0 : ((y == 0.0) ? 1.0 : INFINITY);
    return texp(y * tlog(x));
}
template <int MATH>
__device__ __forceinline__ double ex(double x) { return (MATH == kMathLibm) ? exp(x) : texp(x); }
template <int MATH>
__device__ __forceinline__ double lg(double x) { return (MATH == kMathLibm) ? log(x) : fm::log(x); }
// a / b: IEEE division (with its slow-path branch) in LIBM mode, the branch-free division of soil_math.cuh
// (MUFU seed + Newton + residual correction; normal finite b != 0) in FAST mode
template <int MATH>
__device__ __forceinline__ double dv(double a, double b) { return (MATH == kMathLibm) ? a / b : fm::div(a, b); }

// soil_hydrology_parameterizations.jl:59-63 / :182-186
template <int CLOSURE, int MATH>
__device__ __forceinline__ double matric_potential(const HydroCell &p, double S)
{
    if (CLOSURE == kVanGenuchten) {
        if (MATH == kMathLibm) return -pw<MATH>((pw<MATH>(S, -1.0 / p.m) - 1.0) * pw<MATH>(p.a, -p.b), 1.0 / p.b);
        // (u alpha^-n)^(1/n) = u^(1/n) / alpha
        return -dv<MATH>(pw<MATH>(pw<MATH>(S, -fm::rcp(p.m)) - 1.0, fm::rcp(p.b)), p.a);
    }
    return p.b * pw<MATH>(S, -dv<MATH>(1.0, p.a));
}

// soil_hydrology_parameterizations.jl:72-77 / :195-200 (psi > 0 is an error upstream: NaN here)
template <int CLOSURE, int MATH>
__device__ __forceinline__ double inverse_matric_potential(const HydroCell &p, double psi)
{
    if (psi > 0.0) return NAN;
    if (CLOSURE == kVanGenuchten) return pw<MATH>(1.0 + pw<MATH>(p.a * fabs(psi), p.b), -p.m);
    return pw<MATH>(dv<MATH>(psi, p.b), -p.a);
}

template <int MATH>
__device__ __forceinline__ double effective_saturation(double nu_eff, double theta, double theta_r)
{
    const double theta_safe = fmax(theta, theta_r + kSqrtEps);
    const double nu_safe = fmax(nu_eff, theta_r + kSqrtEps);
    return dv<MATH>(theta_safe - theta_r, nu_safe - theta_r);
}

// soil_heat_parameterizations.jl:35-52; rho_first / rho_second are the positional (_rho_ice, _rho_liq):
// update_aux! passes (rho_l, rho_i), PhaseChange passes (rho_i, rho_l) -- both restated as evaluated there.
template <int CLOSURE, int MATH>
__device__ __forceinline__ double soil_Tf_depressed(const HydroCell &p, double theta_l, double theta_i, double rho_first,
                                                    double rho_second, const ExplicitConst &k, double LH_f0,
                                                    double &psi_w0)
{
    const double theta_tot = fmin(rho_first / rho_second * theta_i + theta_l, p.nu);  // uniform operands: hoisted
    psi_w0 = matric_potential<CLOSURE, MATH>(p, effective_saturation<MATH>(p.nu, theta_tot, p.theta_r));
    return fmax(k.T_freeze * ex<MATH>(dv<MATH>(k.grav * psi_w0, LH_f0)), 1.0);
}

// soil_heat_parameterizations.jl:243-254
template <int MATH>
__device__ __forceinline__ double kappa_sat(double theta_l, double theta_i, double ku, double kf)
{
    const double theta_w = theta_l + theta_i;
    if (theta_w < kEps) return (ku + kf) / 2.0;
    if (MATH == kMathLibm) return pw<MATH>(ku, theta_l / theta_w) * pw<MATH>(kf, theta_i / theta_w);
    if (theta_i == 0.0) return ku;  // ku^1 * kf^0
    const double rw = fm::rcp(theta_w);
    return texp((theta_l * rw) * tlog(ku) + (theta_i * rw) * tlog(kf));
}

// soil_heat_parameterizations.jl:301-323
template <int MATH>
__device__ __forceinline__ double kersten_number(double theta_i, double S_r, double alpha, double beta, double om,
                                                 double quartz, double gravel)
{
    if (theta_i < kEps) {
        if (MATH == kMathLibm)
            return pw<MATH>(S_r, (1.0 + om - alpha * quartz - gravel) / 2.0) *
                   pw<MATH>(pw<MATH>(1.0 + ex<MATH>(-beta * S_r), -3.0) - pw<MATH>((1.0 - S_r) / 2.0, 3.0), 1.0 - om);
        const double e1 = 1.0 + texp(-beta * S_r), h = (1.0 - S_r) / 2.0;
        const double base = fm::rcp(e1 * e1 * e1) - h * h * h;
        // S_r^a * base^b = exp(a log S_r + b log base)
        if (!(base > 0.0)) return (base == 0.0) ? 0.0 : NAN;  // pow(0, b > 0) = 0; a negative base is a DomainError upstream
        return texp(((1.0 + om - alpha * quartz - gravel) / 2.0) * tlog(S_r) + (1.0 - om) * tlog(base));
    }
    return pw<MATH>(S_r, 1.0 + om);
}

// heaviside(x): src/shared_utilities/utils.jl:83-99 (1 if x > eps, else 0)
__device__ __forceinline__ double heaviside(double x) { return (x > kEps) ? 1.0 : 0.0; }

// K (saturation against nu_K) and psi (saturation against nu_psi) of one cell in FAST arithmetic with the table-driven
// log / exp: the same formulas as CellEval<., kMathFast> (soil_closures.cuh; hydraulic_conductivity
// soil_hydrology_parameterizations.jl:161-173 / 239-251, pressure_head :109-127 / 270-289).  When the two porosities
// coincide (no ice) the logarithms and the inner exponential are shared.
template <int CLOSURE>
__device__ __forceinline__ void closure_K_psi_fast(const HydroCell &p, double th, double nu_K, double nu_psi, double &K,
                                                   double &psi)
{
    const double lo = p.theta_r + kSqrtEps;
    const double th_s = fmax(th, lo), num = th_s - p.theta_r;
    const bool shared = (nu_K == nu_psi);
    const double inv_m = (CLOSURE == kVanGenuchten) ? fm::rcp(p.m) : 0.0;
    double L_K = 0.0, E_K = 0.0, l1_K = 0.0;
    {
        const double nu_s = fmax(nu_K, lo), range = nu_s - p.theta_r;
        const double S = fm::div(num, range);
        if (num < range) {
            L_K = tlog(S);
            if (CLOSURE == kVanGenuchten) {
                E_K = L_K * inv_m;
                const double omA = 1.0 - texp(E_K);   // 1 - S^(1/m)
                l1_K = tlog(omA + 1e-300);
                const double t = 1.0 - texp(p.m * l1_K);
                K = (fm::sqrt(S) * (t * t)) * p.K_sat;
            } else {
                K = texp((2.0 * fm::rcp(p.a) + 3.0) * L_K) * p.K_sat;
            }
        } else {
            K = p.K_sat;
        }
    }
    {
        const double nu_s = fmax(nu_psi, lo), range = nu_s - p.theta_r;
        if (num < range) {
            double L = L_K, E = E_K, l1 = l1_K;
            if (!shared) {
                L = tlog(fm::div(num, range));
                if (CLOSURE == kVanGenuchten) {
                    E = L * inv_m;
                    l1 = tlog((1.0 - texp(E)) + 1e-300);
                }
            }
            if (CLOSURE == kVanGenuchten) psi = -(texp((l1 - E) * fm::rcp(p.b)) * fm::rcp(p.a));  // -((S^(-1/m) - 1)^(1/n)) / alpha
            else psi = p.b * texp(-L * fm::rcp(p.a));
        } else {
            const double sat = (th_s - nu_s) * fm::rcp(p.S_s);
            if (CLOSURE == kVanGenuchten) psi = (num == range) ? -0.0 : sat;
            else psi = (num == range) ? p.b : sat + p.b;
        }
    }
}

// AUX: update_aux! for the cell; PHASE: the PhaseChange source of the cell.  With both, the source uses the
// freshly computed theta_l, kappa, T (what the reference reads back from p), so no field is read twice.
template <int CLOSURE, int MATH, bool AUX, bool PHASE>
__global__ void __launch_bounds__(128) k_explicit_cells(const DevView P, const ExplicitView X)
{
    // thread -> cell in the mirrors' memory order, so that a warp's accesses coalesce in either layout:
    // column-fastest: grid (column blocks, levels); level-fastest: a flat grid over ncol * N
    int64_t c;
    int i;
    if (P.sl == 1) {
        const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        c = k / P.N;
        i = (int)(k - c * P.N);
    } else {
        c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        i = blockIdx.y;
    }
    if (c >= P.ncol) return;
    const int64_t q = P.at(i, c);
    const EarthConst &E = P.earth;
    const HydroCell cell = load_cell(P, q);
    const double th = P.Y_theta_l[q], thi = P.Y_theta_i[q];
    const double rcds = __ldg(P.rho_c_ds + q);
    double theta_l, kappa, T, Tf_aux = 0.0, psi_w0_aux = 0.0;
    if (AUX) {
        // theta_l = volumetric_liquid_fraction(theta_l, nu - theta_i, theta_r)
        {
            const double theta_safe = fmax(th, cell.theta_r + kSqrtEps);
            const double nu_safe = fmax(cell.nu - thi, cell.theta_r + kSqrtEps);
            theta_l = (theta_safe < nu_safe) ? theta_safe : nu_safe;
        }
        const double S_r = dv<MATH>(theta_l + thi, cell.nu);
        const double K_e = kersten_number<MATH>(thi, S_r, X.k.alpha, X.k.beta, __ldg(X.nu_ss_om + q),
                                                __ldg(X.nu_ss_quartz + q), __ldg(X.nu_ss_gravel + q));
        const double ks = kappa_sat<MATH>(theta_l, thi, __ldg(X.kappa_sat_unfrozen + q), __ldg(X.kappa_sat_frozen + q));
        kappa = K_e * ks + (1.0 - K_e) * __ldg(X.kappa_dry + q);
        T = E.T_ref + dv<MATH>(P.Y_rho_e[q] + thi * E.rho_i * E.LH_f0, volumetric_heat_capacity(theta_l, thi, rcds, E));
        // K = impedance * viscosity * hydraulic_conductivity(effective_saturation(nu, theta_l(Y), theta_r))
        double Kh, psi, d0, d1;
        if (MATH == kMathFast) {
            closure_K_psi_fast<CLOSURE>(cell, th, cell.nu, cell.nu - thi, Kh, psi);
        } else {
            CellEval<CLOSURE, MATH>(cell, cell.nu).template eval<true, false, false>(th, Kh, d0, d1);
            CellEval<CLOSURE, MATH>(cell, cell.nu - thi).template eval<false, true, false>(th, d0, psi, d1);
        }
        const double f_i = dv<MATH>(thi, theta_l + thi - cell.theta_r);
        const double imp = (MATH == kMathLibm) ? pow(10.0, -X.k.Omega * f_i)
                                               : ((thi == 0.0) ? 1.0 : texp((-X.k.Omega * f_i) * 2.302585092994045684));
        const double visc = ex<MATH>(X.k.gamma * (T - X.k.gammaT_ref));
        X.p_theta_l[q] = theta_l;
        X.p_kappa[q] = kappa;
        X.p_T[q] = T;
        X.p_K[q] = imp * visc * Kh;
        X.p_psi[q] = psi;
        Tf_aux = soil_Tf_depressed<CLOSURE, MATH>(cell, theta_l, thi, E.rho_l, E.rho_i, X.k, E.LH_f0, psi_w0_aux);
        X.p_Tf[q] = Tf_aux;
    } else {
        theta_l = X.p_theta_l[q];
        kappa = X.p_kappa[q];
        T = X.p_T[q];
    }
    if (PHASE) {
        const double dz = P.dz_c[i];
        const double tau = dv<MATH>(3.0 * volumetric_heat_capacity(theta_l, thi, rcds, E) * (dz * dz), kappa);
        double psi_w0, Tf;
        if (AUX && MATH == kMathFast && thi == 0.0) {  // without ice the density ratio does not enter theta_tot
            Tf = Tf_aux;
            psi_w0 = psi_w0_aux;
        } else {
            Tf = soil_Tf_depressed<CLOSURE, MATH>(cell, theta_l, thi, E.rho_i, E.rho_l, X.k, E.LH_f0, psi_w0);
        }
        const double psi_T = E.LH_f0 / X.k.grav * lg<MATH>(dv<MATH>(T, Tf)) * heaviside(Tf - T);
        const double theta_star =
            inverse_matric_potential<CLOSURE, MATH>(cell, psi_w0 + psi_T) * (cell.nu - cell.theta_r) + cell.theta_r;
        const double s = dv<MATH>(theta_l - theta_star, tau);
        if (X.assign_source) {
            X.dYe_theta_l[q] = -s;
            X.dYe_theta_i[q] = (E.rho_l / E.rho_i) * s;
        } else {
            X.dYe_theta_l[q] += -s;
            X.dYe_theta_i[q] += (E.rho_l / E.rho_i) * s;
        }
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

// ---- TOPMODEL runoff -------------------------------------------------------------
// update_infiltration_water_flux!(p, ::TOPMODELRunoff, input, Y, t, model): Runoff/Runoff.jl:234-283;
// topmodel_surface_infiltration :373-376, soil_infiltration_capacity :385-410, topmodel_ss_flux :421-423,
// is_saturated :432-434, update_subsurface_energy_runoff! :266-279.  One thread per column; the three column
// integrals (ice-inclusive and liquid-only saturated thickness, liquid energy of the saturated layers) share one
// sweep over the levels.
struct RunoffView {
    const double *f_max, *precip;               // per column
    double *is_sat, *h_grad, *infiltration, *R_s, *R_ss, *R_ess;
    const double *p_theta_l, *p_T;              // EnergyHydrology: p.soil.theta_l, p.soil.T
    double f_over, R_sb, depth, Omega, gamma, gammaT_ref;
    // clb_soil_step_host: after the runoff of the column (which reads the state at t_n) the integrator's explicit update
    // of the column, u + dt T_exp(u), in place: theta_l, theta_i += dt (PhaseChange source); the infiltration becomes
    // the top water flux of the implicit stage.  apply_dt = 0: none of it.
    double apply_dt = 0.0;
    double *Y_theta_l = nullptr, *Y_theta_i = nullptr, *top_bc_w = nullptr;
    const double *dYe_theta_l = nullptr, *dYe_theta_i = nullptr;
};

template <int MATH>
__global__ void __launch_bounds__(128) k_update_runoff(const DevView P, const RunoffView R)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const bool eh = P.model == 1;
    const EarthConst &E = P.earth;
    double h_all = 0.0, h_liq = 0.0, e_liq = 0.0;
    for (int i = 0; i < P.N; ++i) {
        const int64_t q = P.at(i, c);
        const double th = P.Y_theta_l[q], thi = eh ? P.Y_theta_i[q] : 0.0;
        const double nu = __ldg(P.nu + q), theta_r = __ldg(P.theta_r + q);
        const double range = nu - theta_r, dz = P.dz_c[i];
        const double s_all = heaviside((th + thi - theta_r) - range) * (th + thi - theta_r) / range;
        const double s_liq = heaviside((th - theta_r) - range) * (th - theta_r) / range;
        R.is_sat[q] = s_liq;
        h_all += s_all * dz;
        h_liq += s_liq * dz;
        if (eh) e_liq += s_liq * volumetric_internal_energy_liq(R.p_T[q], E) * dz;
    }
    const int64_t qt = P.at(P.N - 1, c);
    double ic = -1 * __ldg(P.K_sat + qt);
    if (eh) {
        const double thi = P.Y_theta_i[qt];
        const double f_i = thi / (R.p_theta_l[qt] + thi - __ldg(P.theta_r + qt));
        const double imp = (MATH == kMathLibm) ? pow(10.0, -R.Omega * f_i) : texp((-R.Omega * f_i) * 2.302585092994045684);
        ic = -__ldg(P.K_sat + qt) * imp * ex<MATH>(R.gamma * (R.p_T[qt] - R.gammaT_ref));
    }
    const double precip = R.precip[c];
    const double f_sat = fmin(R.f_max[c] * ex<MATH>(-R.f_over / 2.0 * (R.depth - h_all)), 1.0);
    const double inf = (1.0 - f_sat) * fmax(ic, precip);
    const double R_ss = R.R_sb * ex<MATH>(-R.f_over * (R.depth - h_liq));
    R.infiltration[c] = inf;
    R.R_s[c] = fabs(precip - inf);
    R.h_grad[c] = h_liq;
    R.R_ss[c] = R_ss;
    if (eh) R.R_ess[c] = e_liq * (R_ss / fmax(h_liq, kEps));
    if (R.apply_dt != 0.0) {
        R.top_bc_w[c] = inf;
        for (int i = 0; i < P.N; ++i) {  // product and sum rounded separately, as `@. u + dt * du` is
            const int64_t q = P.at(i, c);
            R.Y_theta_l[q] = __dadd_rn(R.Y_theta_l[q], __dmul_rn(R.apply_dt, R.dYe_theta_l[q]));
            R.Y_theta_i[q] = __dadd_rn(R.Y_theta_i[q], __dmul_rn(R.apply_dt, R.dYe_theta_i[q]));
        }
    }
}

}  // namespace clb
