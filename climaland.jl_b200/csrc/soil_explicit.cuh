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
    // quotients of launch constants, divided on the host (make_explicit_view: the same IEEE quotient, i.e. the same bits):
    // in k_explicit_cells_uniform every `/` of two constants was an inlined IEEE division per THREAD -- four of them on
    // a cell's path, 155 of the kernel's ~3 200 issue cycles (tools/sass_cost_lines.py)
    double rho_l_over_rho_i = 0.0, rho_i_over_rho_l = 0.0, LH_f0_over_grav = 0.0, inv_LH_f0 = 0.0, inv_rho_l = 0.0;
};

// per-cell fields only the explicit stage touches
struct ExplicitView {
    const double *kappa_dry, *kappa_sat_unfrozen, *kappa_sat_frozen, *nu_ss_om, *nu_ss_quartz, *nu_ss_gravel;
    double *p_theta_l, *p_kappa, *p_K, *p_T, *p_psi, *p_Tf;  // p.soil.{theta_l, kappa, K, T, psi, Tf_depressed}
    double *total_water, *total_energy;                     // per column
    double *dYe_theta_l, *dYe_theta_i;                      // explicit tendency the PhaseChange source adds into
    ExplicitConst k;
    int assign_source = 0;                                  // 1: the source is stored, not added (clb_soil_step_host)
    // k_explicit_cells_uniform, clb_soil_step: != 0: the integrator's explicit update of the cell, u + dt T_exp(u), in
    // place (Y_theta_l, Y_theta_i += dt * source, product and sum rounded separately as `@. u + dt * du` is) instead
    // of a stored source -- the per-column sweep, which needs the state at t_n, has then already run
    double apply_dt = 0.0;
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

// the per-column sweep of the explicit stage (k_update_runoff below; the SWEEP form of k_explicit_cells_uniform)
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
    // the column integrals of update_aux! (k_explicit_totals) taken in the same sweep over the levels; nullptr: not here
    double *total_water = nullptr, *total_energy = nullptr;
    int model = 2;  // CLB_RUNOFF_*: 0 NoRunoff, 1 SurfaceRunoff, 2 TOPMODELRunoff (k_update_runoff only)
};

// ---- FAST mode, warp-uniform control flow ------------------------------------------------------------------
// The kernel above follows the reference's case distinctions cell by cell: frozen / unfrozen (Kersten number, kappa_sat,
// the shared closure, the second depressed freezing point), saturated / unsaturated, T above / below T_f.  With the
// cases mixed inside a warp its lanes idle on each other's paths (ncu, profiles/r1: 19 of 32 lanes active, 45.5 M
// warp instructions per ~1 degree domain -- more than the implicit stage).  Here every decision that changes the
// INSTRUCTION STREAM is taken per warp (__any_sync votes), and per lane only as a select between finished values:
//   * no lane of the warp holds ice: the no-ice shortcuts for the whole warp (what a summer warp of a real domain,
//     whose ice is spatially coherent, takes); otherwise the general formulas for every lane -- they reduce to the
//     shortcut's VALUE where theta_i = 0 (same saturation, same theta_tot, 10^0 = 1) -- so a mixed warp costs the
//     general path once instead of both paths one after the other;
//   * saturated cells run the unsaturated formulas on finite garbage and select, as the lane kernels' closure does;
//   * the logarithm of T / T_f (series log: it needs relative accuracy near 1) only where some lane is below T_f.
// Every value is the one the kernel above produces (same formulas and functions) up to the last bit or two: kappa_sat
// of an ice-free cell in a warp with ice is exp(log k_u) instead of k_u (<= 2^-52 relative), and the series of the
// table-driven log / exp are summed in Horner order here (xbf::Tab).
constexpr int kExplicitStage = 17;  // per-cell inputs of k_explicit_cells_uniform staged in shared memory

namespace xbf {

// The constants of the table-driven log / exp in CONSTANT memory: an FP64 instruction takes one operand straight from
// a constant bank, but a 64-bit literal that is not a bank operand costs two moves (IMAD.MOV / UMOV of its halves)
// EVERY time it is used -- with one cell per thread nothing amortises them (the lane kernels use a literal for W = 4
// chains): ncu's opcode mix of this kernel had 230 such moves among 1 811 instructions per warp.
__constant__ double c_xm[12] = {mtab::kLn2, mtab::kExpScale, -mtab::kLn2NHi, -mtab::kLn2NLo, mtab::kExpC2, mtab::kExpC3,
                                mtab::kExpC4, mtab::kExpC5, 0.33333333333333333, 0.2, -0.16666666666666667,
                                4503601774854144.0};

// tlog / texp with the tables in SHARED memory (2.5 KB per block, filled by its 128 threads): a look-up through L1
// from global memory stalls the warp on the long scoreboard (ncu: 3.8 stall cycles per issued instruction, the
// largest item), from shared memory it is a ~25-cycle access.  The algorithms of fmv::log_tab<1> / exp_tab<1>
// with the constants from constant memory and the series in Horner form (values within an ulp of theirs).
struct Tab {
    fmv::MathTab MT;
    __device__ __forceinline__ double log(double x) const
    {
        const int hx = __double2hiint(x);
        const int tmp = hx - mtab::kLogOffHi;
        const int k = tmp >> 20;
        const double2 e = MT.logt[(tmp >> 13) & (mtab::kLogN - 1)];
        const double z = __hiloint2double(hx - (tmp & 0xfff00000), __double2loint(x));
        const double dk = __int2double_rn(k);
        double r = fma(z, e.x, -1.0);
        const double w = fma(dk, c_xm[0], e.y);
        const double r2 = r * r;
        // Horner in r (the lane kernels' Estrin form needs a second constant in a register per pair of coefficients --
        // one LDC each here, where a thread holds one cell -- and two more DFMAs with three live register operands):
        // the same series, rounded in another order (both within the 2^-56 max(1, |log x|) of tests/test_cuda_math.py);
        // static issue cost of the whole kernel -1.7 %, whole soil step 125.15 -> 124.4 us
        double p01 = fma(r, c_xm[10], c_xm[9]);
        p01 = fma(r, p01, -0.25);
        p01 = fma(r, p01, c_xm[8]);
        p01 = fma(r, p01, -0.5);
        r = fma(r2, p01, r);
        return w + r;
    }
    __device__ __forceinline__ double exp(double x) const
    {
        const double MAGIC = 6755399441055744.0;
        double t = fma(x, c_xm[1], MAGIC);
        const int k = __double2loint(t);
        t = t - MAGIC;
        const double tb = MT.expt[k & (mtab::kExpN - 1)];
        double r = fma(t, c_xm[2], x);
        r = fma(t, c_xm[3], r);
        const double r2 = r * r;
        double q23 = fma(r, c_xm[7], c_xm[6]);  // Horner, as in log above
        q23 = fma(r, q23, c_xm[5]);
        q23 = fma(r, q23, c_xm[4]);
        r = fma(r2, q23, r);
        r = fma(tb, r, tb);
        const int kc = min(max(k >> 6, -1021), 1022);
        return __hiloint2double(__double2hiint(r) + (kc << 20), __double2loint(r));
    }
};

// van Genuchten / Brooks-Corey matric potential at saturation S in (0, 1] (soil_hydrology_parameterizations.jl:59-63,
// 182-186), with the reciprocals of m, n (or of c) passed in
template <int CLOSURE>
__device__ __forceinline__ double matric_potential(const Tab &M, const HydroCell &p, double inv_m, double inv_n, double S)
{
    if (CLOSURE == kVanGenuchten) {
        const double u = M.exp(-inv_m * M.log(S)) - 1.0;            // S^(-1/m) - 1 >= 0
        const double v = M.exp(inv_n * M.log(u + 1e-300));          // u = 0: selected below
        return -(((u == 0.0) ? 0.0 : v) * fm::rcp(p.a));         // the closure's own 1/alpha (one rcp per cell after CSE)
    }
    return p.b * M.exp(-inv_m * M.log(S));                           // inv_m carries 1/c
}

template <int CLOSURE>
__device__ __forceinline__ double inverse_matric_potential(const Tab &M, const HydroCell &p, double psi)
{
    double r;
    if (CLOSURE == kVanGenuchten) {
        const double x = p.a * fabs(psi);
        const double inner = M.exp(p.b * M.log(x + 1e-300));         // (alpha |psi|)^n; x = 0: 0
        r = M.exp(-p.m * M.log(1.0 + ((x == 0.0) ? 0.0 : inner)));
    } else {
        const double x = fm::div(psi, p.b);                        // psi / psi_b >= 0
        r = (x == 0.0) ? INFINITY : M.exp(-p.a * M.log(x + 1e-300));
    }
    return (psi > 0.0) ? NAN : r;
}

template <int CLOSURE>
__device__ __forceinline__ double Tf_depressed(const Tab &M, const HydroCell &p, double inv_m, double inv_n, double theta_l, double theta_i,
                                               double ratio, const ExplicitConst &k, double LH_f0, double &psi_w0)
{
    const double theta_tot = fmin(ratio * theta_i + theta_l, p.nu);
    const double lo = p.theta_r + kSqrtEps;
    const double S = fm::div(fmax(theta_tot, lo) - p.theta_r, fmax(p.nu, lo) - p.theta_r);
    psi_w0 = matric_potential<CLOSURE>(M, p, inv_m, inv_n, S);
    return fmax(k.T_freeze * M.exp((k.grav * psi_w0) * k.inv_LH_f0), 1.0);  // |g psi / LH| ~ 1e-3: an ulp of it is nothing
}

}  // namespace xbf

// SWEEP (column-fastest mirrors, N <= 32; a block is COLS columns x N levels, rounded up to whole warps): the per-column sweep of
// the explicit stage -- TOPMODEL runoff (k_update_runoff) and the column integrals of update_aux! -- in the SAME
// kernel: every cell leaves its five dz-weighted terms in shared memory, and the warp of the top level, which holds the
// top cell's T and theta_l in registers, adds them up level by level (the order of the sweep kernel: same bits) and
// finishes the column.  Each cell then applies its own explicit update in place (X.apply_dt), so a whole explicit stage
// is ONE launch that reads every field once.
template <int CLOSURE, bool AUX, bool PHASE, bool SWEEP = false, int COLS = 32>
#ifndef CLB_XBF_MINB
#define CLB_XBF_MINB 8  // 64 registers (a 52-byte spill) against 72 at 7 blocks per SM
#endif
__global__ void __launch_bounds__(SWEEP ? 32 * COLS : 128, SWEEP ? 32 / COLS : CLB_XBF_MINB)  // 64 registers either way
    k_explicit_cells_uniform(const DevView P, const ExplicitView X, const RunoffView R)
{
    constexpr unsigned kFull = 0xffffffffu;
    static_assert(mtab::kLogN <= 128 && mtab::kExpN <= 128, "one table entry per thread of the block");
    static_assert(!SWEEP || (AUX && PHASE), "the sweep belongs to the whole explicit stage");
    __shared__ __align__(16) unsigned char tab_sm[fmv::kMathTabBytes];
    // dynamic shared memory: [kExplicitStage fields][blockDim] staging of the cell's inputs, then (SWEEP) the column
    // terms [5][N levels][COLS columns]
    extern __shared__ double dyn_sm[];
    double *const stage = dyn_sm + threadIdx.x;
    const int nt = blockDim.x;
    double *const sweep_sm = dyn_sm + (size_t)kExplicitStage * nt;
    int64_t c;
    int i;
    bool pad = false;  // SWEEP: threads that round the block up to whole warps
    if (SWEEP) {
        c = (int64_t)blockIdx.x * COLS + (threadIdx.x % COLS);
        i = threadIdx.x / COLS;
        pad = i >= P.N;
        i = pad ? P.N - 1 : i;
    } else if (P.sl == 1) {
        int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t last = P.ncol * P.N - 1;
        k = (k > last) ? -1 : k;
        c = (k < 0) ? P.ncol : k / P.N;
        i = (k < 0) ? 0 : (int)(k - c * P.N);
    } else {
        c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        i = blockIdx.y;
    }
    // every lane stays alive for the warp votes: lanes past the last column work on a copy of it and store nothing
    const bool live = c < P.ncol && !pad;
    if (c >= P.ncol) c = P.ncol - 1;
    const int64_t q = P.at(i, c);
    const EarthConst &E = P.earth;
    // Every per-cell input goes HBM -> shared memory by cp.async, all requests in flight before anything waits: read
    // one by one where they are used (the registers do not hold 17 doubles ahead of time), the loads were ~5 dependent
    // HBM round trips per cell (ncu: long_scoreboard 3.2 stall cycles per issued instruction, the largest item).  A
    // thread only reads what it staged itself: no block barrier.
    {
        // every field is there (explicit_ready / clb_soil_step require them) except m of a Brooks-Corey handle: the
        // pointers are not tested one by one (17 uniform null checks were 54 instructions per thread)
        auto cpa = [&](int slot, const double *src) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + (size_t)slot * nt);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + q) : "memory");
        };
        // (Programmatic dependent launch of this kernel and an early release of the implicit stage behind it were
        // measured: 130.8 against 127.3 us per whole step -- the stage's blocks, which need a whole SM's shared memory,
        // get in the way of this grid's last wave -- and not kept.  So were PERSISTENT blocks walking over the tiles, to
        // fill the tables and wait for them once per block instead of once per tile (that barrier holds 12 % of the
        // stall samples): 147.3 us -- without a second staging buffer, which does not fit, every tile's HBM latency is
        // then exposed to its block instead of being covered by the next block the hardware schedules.)
        cpa(0, P.nu); cpa(1, P.theta_r); cpa(2, P.K_sat); cpa(3, P.S_s); cpa(4, P.hcm_a); cpa(5, P.hcm_b);
        if (CLOSURE == kVanGenuchten) cpa(6, P.hcm_m);
        cpa(7, P.Y_theta_l); cpa(8, P.Y_theta_i); cpa(9, P.rho_c_ds); cpa(10, P.Y_rho_e);
        if (AUX) {
            cpa(11, X.nu_ss_om); cpa(12, X.nu_ss_quartz); cpa(13, X.nu_ss_gravel); cpa(14, X.kappa_sat_unfrozen);
            cpa(15, X.kappa_sat_frozen); cpa(16, X.kappa_dry);
        } else {
            cpa(11, X.p_theta_l); cpa(12, X.p_kappa); cpa(13, X.p_T);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // (Measured and not kept: prefetch.global.L2 of the inputs of the block that takes this block's place on the SM,
        // 148 / 296 / 592 blocks ahead, one lane per 128-byte line: 128.9 / 130.5 / 134.1 against 127.5 us per whole soil
        // step -- the wait of a fresh block is not what the prefetch can remove, and its 17 requests per line cost more.)
    }
    // not unrolled: all but the first 128 threads of a block only fall through (unrolled, the loop was 85 instructions
    // for every thread)
#pragma unroll 1
    for (int t = threadIdx.x; t < mtab::kLogN; t += blockDim.x) fmv::math_tab_fill(tab_sm, t);
    const xbf::Tab M{fmv::MathTab{reinterpret_cast<const double2 *>(tab_sm),
                                  reinterpret_cast<const double *>(tab_sm + mtab::kLogN * 16)}};
    __syncthreads();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    auto in = [&](int slot) { return stage[(size_t)slot * nt]; };
    HydroCell cell;
    cell.nu = in(0); cell.theta_r = in(1); cell.K_sat = in(2); cell.S_s = in(3); cell.a = in(4); cell.b = in(5);
    cell.m = (CLOSURE == kVanGenuchten) ? in(6) : 0.0;
    const double th = in(7), thi = in(8);
    const double rcds = in(9);
    const bool any_ice = __any_sync(kFull, thi != 0.0);
    const double inv_m = (CLOSURE == kVanGenuchten) ? fm::rcp(cell.m) : fm::div(1.0, cell.a);
    const double inv_n = (CLOSURE == kVanGenuchten) ? fm::rcp(cell.b) : 0.0;
    const double lo = cell.theta_r + kSqrtEps;
    double theta_l, kappa, T, Tf_aux = 0.0, psi_w0_aux = 0.0, rho_e = 0.0;
    if (AUX) {
        const double th_s = fmax(th, lo), num = th_s - cell.theta_r;
        const double nu_psi = fmax(cell.nu - thi, lo), nu_K = fmax(cell.nu, lo);
        theta_l = (th_s < nu_psi) ? th_s : nu_psi;
        const double S_r = fm::div(theta_l + thi, cell.nu);
        // ---- Kersten number (soil_heat_parameterizations.jl:301-323): both forms share log S_r
        const double om = in(11);
        const bool unfrozen = thi < kEps;
        const double lSr = M.log(S_r);
        double K_e = 0.0;
        if (__any_sync(kFull, unfrozen)) {
            const double e1 = 1.0 + M.exp(-X.k.beta * S_r), h = (1.0 - S_r) / 2.0;
            const double base = fm::rcp(e1 * e1 * e1) - h * h * h;
            const double v = M.exp(((1.0 + om - X.k.alpha * in(12) - in(13)) / 2.0) * lSr +
                                  (1.0 - om) * M.log(fmax(base, 1e-300)));
            K_e = (base > 0.0) ? v : ((base == 0.0) ? 0.0 : NAN);
        }
        if (__any_sync(kFull, !unfrozen)) {
            const double v = M.exp((1.0 + om) * lSr);
            K_e = unfrozen ? K_e : v;
        }
        // ---- kappa_sat (:243-254)
        const double ku = in(14), kf = in(15);
        double ks = ku;
        if (any_ice) {
            const double theta_w = theta_l + thi;
            const double rw = fm::rcp(theta_w);
            const double v = M.exp((theta_l * rw) * M.log(ku) + (thi * rw) * M.log(kf));
            ks = (thi == 0.0) ? ku : v;
            ks = (theta_w < kEps) ? (ku + kf) / 2.0 : ks;
        }
        kappa = K_e * ks + (1.0 - K_e) * in(16);
        rho_e = in(10);
        T = E.T_ref + fm::div(rho_e + thi * E.rho_i * E.LH_f0, volumetric_heat_capacity(theta_l, thi, rcds, E));
        // ---- K at the saturation against nu, psi at the saturation against nu - theta_i (closure_K_psi_fast)
        double Kh, psi;
        {
            const double range_K = nu_K - cell.theta_r, range_p = nu_psi - cell.theta_r;
            const double S_K = fm::div(num, range_K);
            const double L_K = M.log(S_K);
            double L_p = L_K, E_K = 0.0, l1_K = 0.0;
            if (CLOSURE == kVanGenuchten) {
                E_K = L_K * inv_m;
                const double omA = 1.0 - M.exp(E_K);
                l1_K = M.log(fmax(omA, 0.0) + 1e-300);
                const double t = 1.0 - M.exp(cell.m * l1_K);
                Kh = (fm::sqrt(S_K) * (t * t)) * cell.K_sat;
            } else {
                Kh = M.exp((2.0 * inv_m + 3.0) * L_K) * cell.K_sat;
            }
            Kh = (num < range_K) ? Kh : cell.K_sat;
            double E_p = E_K, l1_p = l1_K;
            if (any_ice) {
                L_p = M.log(fm::div(num, range_p));
                if (CLOSURE == kVanGenuchten) {
                    E_p = L_p * inv_m;
                    l1_p = M.log(fmax(1.0 - M.exp(E_p), 0.0) + 1e-300);
                }
                // a lane without ice: range_p == range_K, the same operands, the same values
            }
            const double sat = (th_s - nu_psi) * fm::rcp(cell.S_s);
            if (CLOSURE == kVanGenuchten) {
                const double un = -(M.exp((l1_p - E_p) * inv_n) * fm::rcp(cell.a));
                psi = (num < range_p) ? un : ((num == range_p) ? -0.0 : sat);
            } else {
                const double un = cell.b * M.exp(-L_p * inv_m);
                psi = (num < range_p) ? un : ((num == range_p) ? cell.b : sat + cell.b);
            }
        }
        const double f_i = fm::div(thi, theta_l + thi - cell.theta_r);
        const double imp = any_ice ? M.exp((-X.k.Omega * f_i) * 2.302585092994045684) : 1.0;  // 10^0 = exp(0) = 1 exactly
        const double visc = M.exp(X.k.gamma * (T - X.k.gammaT_ref));
        Tf_aux = xbf::Tf_depressed<CLOSURE>(M, cell, inv_m, inv_n, theta_l, thi, X.k.rho_l_over_rho_i, X.k, E.LH_f0, psi_w0_aux);
        if (live) {
            X.p_theta_l[q] = theta_l;
            X.p_kappa[q] = kappa;
            X.p_T[q] = T;
            X.p_K[q] = imp * visc * Kh;
            X.p_psi[q] = psi;
            X.p_Tf[q] = Tf_aux;
        }
    } else {
        theta_l = in(11);
        kappa = in(12);
        T = in(13);
    }
    if (PHASE) {
        const double dz = P.dz_c[i];
        const double tau = fm::div(3.0 * volumetric_heat_capacity(theta_l, thi, rcds, E) * (dz * dz), kappa);
        double psi_w0 = psi_w0_aux, Tf = Tf_aux;
        if (!AUX || any_ice)  // without ice the density ratio does not enter theta_tot: the value of update_aux!
            Tf = xbf::Tf_depressed<CLOSURE>(M, cell, inv_m, inv_n, theta_l, thi, X.k.rho_i_over_rho_l, X.k, E.LH_f0, psi_w0);
        const bool below = (Tf - T) > kEps;  // heaviside(Tf - T)
        double psi_T = 0.0;
        if (__any_sync(kFull, below)) psi_T = below ? X.k.LH_f0_over_grav * fm::log(fm::div(T, Tf)) : 0.0;
        const double theta_star =
            xbf::inverse_matric_potential<CLOSURE>(M, cell, psi_w0 + psi_T) * (cell.nu - cell.theta_r) + cell.theta_r;
        const double s = fm::div(theta_l - theta_star, tau);
        if (live) {
            if (X.apply_dt != 0.0) {
                P.Y_theta_l[q] = __dadd_rn(th, __dmul_rn(X.apply_dt, -s));
                P.Y_theta_i[q] = __dadd_rn(thi, __dmul_rn(X.apply_dt, X.k.rho_l_over_rho_i * s));
            } else if (X.assign_source) {
                X.dYe_theta_l[q] = -s;
                X.dYe_theta_i[q] = X.k.rho_l_over_rho_i * s;
            } else {
                X.dYe_theta_l[q] += -s;
                X.dYe_theta_i[q] += X.k.rho_l_over_rho_i * s;
            }
        }
    }
    if (SWEEP) {
        // update_infiltration_water_flux!(p, ::TOPMODELRunoff, ...) (Runoff/Runoff.jl:234-283) and the column integrals
        // of update_aux! (energy_hydrology.jl:1282-1327); expressions and summation order of k_update_runoff
        const int N = P.N;
        const double dz = P.dz_c[i], range = cell.nu - cell.theta_r;
        const double a_all = th + thi - cell.theta_r, a_liq = th - cell.theta_r;
        const bool sat_all = (a_all - range) > kEps, sat_liq = (a_liq - range) > kEps;  // heaviside
        double s_all = 0.0, s_liq = 0.0;
        if (__any_sync(kFull, sat_all)) {  // the division only where some column of the warp is saturated at this level
            s_all = sat_all ? a_all / range : 0.0;
            s_liq = sat_liq ? a_liq / range : 0.0;
        }
        const int TS = N * COLS;
        if (!pad) {
            double *const term = sweep_sm + (size_t)i * COLS + (threadIdx.x % COLS);
            term[0] = s_all * dz;
            term[TS] = s_liq * dz;
            term[2 * TS] = s_liq * volumetric_internal_energy_liq(T, E) * dz;
            term[3 * TS] = (th + fm::div_by(thi * E.rho_i, E.rho_l, X.k.inv_rho_l)) * dz;  // the bits of (thi rho_i) / rho_l
            term[4 * TS] = rho_e * dz;
        }
        if (live) R.is_sat[q] = s_liq;
        __syncthreads();
        if (i == N - 1 && !pad) {  // the top level's threads: one per column
            double h_all = 0.0, h_liq = 0.0, e_liq = 0.0, tw = 0.0, te = 0.0;
            const double *col = sweep_sm + (threadIdx.x % COLS);
            for (int l = 0; l < N; ++l) {
                h_all += col[l * COLS];
                h_liq += col[TS + l * COLS];
                e_liq += col[2 * TS + l * COLS];
                tw += col[3 * TS + l * COLS];
                te += col[4 * TS + l * COLS];
            }
            const double f_i = thi / (theta_l + thi - cell.theta_r);
            const double imp = texp((-R.Omega * f_i) * 2.302585092994045684);
            const double ic = -cell.K_sat * imp * texp(R.gamma * (T - R.gammaT_ref));
            const double precip = R.precip[c];
            const double f_sat = fmin(R.f_max[c] * texp(-R.f_over / 2.0 * (R.depth - h_all)), 1.0);
            const double inf = (1.0 - f_sat) * fmax(ic, precip);
            const double R_ss = R.R_sb * texp(-R.f_over * (R.depth - h_liq));
            if (live) {
                R.infiltration[c] = inf;
                R.R_s[c] = fabs(precip - inf);
                R.h_grad[c] = h_liq;
                R.R_ss[c] = R_ss;
                R.R_ess[c] = e_liq * (R_ss / fmax(h_liq, kEps));
                R.total_water[c] = tw;
                R.total_energy[c] = te;
                if (R.top_bc_w) R.top_bc_w[c] = inf;
            }
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

template <int MATH>
__global__ void __launch_bounds__(128) k_update_runoff(const DevView P, const RunoffView R)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const bool eh = P.model == 1;
    const EarthConst &E = P.earth;
    double h_all = 0.0, h_liq = 0.0, e_liq = 0.0, tw = 0.0, te = 0.0;
    const bool totals = R.total_water != nullptr;
    // R.p_T == nullptr (the sweep runs BEFORE update_aux!, clb_soil_step): p.soil.T and p.soil.theta_l of the cell are
    // evaluated here with update_aux!'s own expressions (energy_hydrology.jl:745-775), i.e. to the same bits
    const bool own_T = eh && R.p_T == nullptr;
    // CH levels at a time, every load of the chunk issued before the first use: a thread owns a whole column, so the
    // memory-level parallelism of this HBM-bound sweep is what the chunk holds in flight
    constexpr int CH = 5;
    double T_top = 0.0, thl_top = 0.0;
    for (int i0 = 0; i0 < P.N; i0 += CH) {
        double th[CH], thi[CH], nu[CH], theta_r[CH], Tq[CH], rho_e[CH], rcds[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int i = min(i0 + j, P.N - 1);
            const int64_t q = P.at(i, c);
            th[j] = P.Y_theta_l[q];
            thi[j] = eh ? P.Y_theta_i[q] : 0.0;
            nu[j] = __ldg(P.nu + q);
            theta_r[j] = __ldg(P.theta_r + q);
            Tq[j] = (eh && !own_T) ? R.p_T[q] : 0.0;
            rho_e[j] = (eh && (own_T || totals)) ? P.Y_rho_e[q] : 0.0;
            rcds[j] = own_T ? __ldg(P.rho_c_ds + q) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int i = i0 + j;
            if (i < P.N) {
                const int64_t q = P.at(i, c);
                const double range = nu[j] - theta_r[j], dz = P.dz_c[i];
                if (totals) {  // total_liq_water_vol_per_area!, total_energy_per_area! (energy_hydrology.jl:1282-1327)
                    tw += (th[j] + thi[j] * E.rho_i / E.rho_l) * dz;
                    te += rho_e[j] * dz;
                }
                double T = Tq[j];
                if (own_T) {
                    const double lo = theta_r[j] + kSqrtEps;
                    const double th_s = fmax(th[j], lo), nu_s = fmax(nu[j] - thi[j], lo);
                    const double theta_l = (th_s < nu_s) ? th_s : nu_s;
                    T = E.T_ref + dv<MATH>(rho_e[j] + thi[j] * E.rho_i * E.LH_f0, volumetric_heat_capacity(theta_l, thi[j], rcds[j], E));
                    if (i == P.N - 1) { T_top = T; thl_top = theta_l; }
                }
                if (R.model == 2) {
                    const double s_all = heaviside((th[j] + thi[j] - theta_r[j]) - range) * (th[j] + thi[j] - theta_r[j]) / range;
                    const double s_liq = heaviside((th[j] - theta_r[j]) - range) * (th[j] - theta_r[j]) / range;
                    R.is_sat[q] = s_liq;
                    h_all += s_all * dz;
                    h_liq += s_liq * dz;
                    if (eh) e_liq += s_liq * volumetric_internal_energy_liq(T, E) * dz;
                } else if (R.model == 1) {  // SurfaceRunoff: is_saturated(theta_l + theta_i, nu), Runoff.jl:139, 432-434
                    const double s = heaviside((th[j] + thi[j]) - nu[j]);
                    R.is_sat[q] = s;
                    if (i == P.N - 1) h_all = s;  // the top centre's value (top_center_to_surface)
                }
            }
        }
    }
    const int64_t qt = P.at(P.N - 1, c);
    double ic = -1 * __ldg(P.K_sat + qt);
    if (eh) {
        const double thi = P.Y_theta_i[qt];
        const double thl = own_T ? thl_top : R.p_theta_l[qt], Tt = own_T ? T_top : R.p_T[qt];
        const double f_i = thi / (thl + thi - __ldg(P.theta_r + qt));
        const double imp = (MATH == kMathLibm) ? pow(10.0, -R.Omega * f_i) : texp((-R.Omega * f_i) * 2.302585092994045684);
        ic = -__ldg(P.K_sat + qt) * imp * ex<MATH>(R.gamma * (Tt - R.gammaT_ref));
    }
    const double precip = R.precip[c];
    double inf;
    if (R.model == 2) {
        const double f_sat = fmin(R.f_max[c] * ex<MATH>(-R.f_over / 2.0 * (R.depth - h_all)), 1.0);
        inf = (1.0 - f_sat) * fmax(ic, precip);
        const double R_ss = R.R_sb * ex<MATH>(-R.f_over * (R.depth - h_liq));
        R.h_grad[c] = h_liq;
        R.R_ss[c] = R_ss;
        if (eh) R.R_ess[c] = e_liq * (R_ss / fmax(h_liq, kEps));
    } else if (R.model == 1) {
        inf = (1 - h_all) * fmax(ic, precip);  // surface_infiltration, Runoff.jl:109-111
    } else {
        inf = precip;                          // NoRunoff, Runoff.jl:69-71
    }
    R.infiltration[c] = inf;
    if (R.model != 0) R.R_s[c] = fabs(precip - inf);
    if (totals) {
        R.total_water[c] = tw;
        R.total_energy[c] = te;
    }
    if (R.top_bc_w) R.top_bc_w[c] = inf;
    if (R.apply_dt != 0.0) {
        for (int i0 = 0; i0 < P.N; i0 += CH) {  // product and sum rounded separately, as `@. u + dt * du` is
            double yl[CH], yi[CH], dl[CH], di[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int64_t q = P.at(min(i0 + j, P.N - 1), c);
                yl[j] = R.Y_theta_l[q]; yi[j] = R.Y_theta_i[q]; dl[j] = R.dYe_theta_l[q]; di[j] = R.dYe_theta_i[q];
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (i0 + j < P.N) {
                    const int64_t q = P.at(i0 + j, c);
                    R.Y_theta_l[q] = __dadd_rn(yl[j], __dmul_rn(R.apply_dt, dl[j]));
                    R.Y_theta_i[q] = __dadd_rn(yi[j], __dmul_rn(R.apply_dt, di[j]));
                }
            }
        }
    }
}

// soil_boundary_fluxes!(bc::AtmosDrivenFluxBC, Val((:soil,)), ...) after the runoff: boundary_conditions.jl:922-935,
// 988-1002.  heat == nullptr (RichardsModel): top_bc = infiltration (boundary_flux! :200-213).
struct AtmosView {
    const double *infiltration, *vapor_flux_liq, *lhf, *shf, *R_n, *T_air;
    double *top_bc_w, *top_bc_h;
};
__global__ void __launch_bounds__(128) k_atmos_driven_fluxes(const DevView P, const AtmosView A)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const double inf = A.infiltration[c];
    if (!A.top_bc_h) {
        A.top_bc_w[c] = inf;
        return;
    }
    A.top_bc_w[c] = inf + A.vapor_flux_liq[c];
    A.top_bc_h[c] = A.R_n[c] + A.lhf[c] + A.shf[c] + inf * volumetric_internal_energy_liq(A.T_air[c], P.earth);
}

// soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...): boundary_conditions.jl:590-608
__global__ void __launch_bounds__(128) k_energy_water_free_drainage(const DevView P, const double *p_K, const double *p_T,
                                                                     double *bot_bc_w, double *bot_bc_h)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const int64_t q = P.at(0, c);
    const double K = p_K[q];
    bot_bc_w[c] = -1 * K;
    bot_bc_h[c] = -1 * K * volumetric_internal_energy_liq(p_T[q], P.earth);
}

}  // namespace clb
