// soil_warp.cuh -- lane-per-cell variant of the fused implicit stage.
//
// One column occupies one SEG-lane segment of a warp (SEG = 16: two columns per warp,
// SEG = 32: one), lane l holds level l: its parameters, lagged cache and state stay in
// that lane's registers for the whole Newton loop, so every field is read from HBM once
// and written once.  The closures (the FP64-heavy part) run on all levels at once, the
// stencil exchanges neighbours with warp shuffles (face fluxes are computed once, by the
// lane below the face, and shuffled up, so the divergence telescopes exactly), and the
// tridiagonal systems are solved by parallel cyclic reduction across the segment
// (log2(SEG) shuffle steps).  With the level-fastest mirror layout (sl = 1, sc = N: the
// reference's own layout) a warp's loads are contiguous.
//
// Compared with one thread per column this does ~1.5x the FP64 work per column (PCR, idle
// pad lanes) but has SEG x the parallelism and a small register footprint, which is what
// the ~1 degree global column count (61 206) needs: one thread per column cannot fill
// 148 SMs there (profiles/).
//
// W22 = (rho_e_int, rho_e_int) of EnergyHydrology depends only on lagged fields
// (kappa, theta_l, theta_i: energy_hydrology.jl:559-573), so its cyclic reduction
// coefficients are computed once per stage and re-applied to each Newton right-hand side.
#pragma once
#include "soil_device.cuh"
#include "soil_fused.cuh"
#include "soil_hooks.cuh"

namespace clb {

constexpr unsigned kFull = 0xffffffffu;

template <int SEG>
__device__ __forceinline__ double from_above(double v, int d = 1) { return __shfl_down_sync(kFull, v, d, SEG); }
template <int SEG>
__device__ __forceinline__ double from_below(double v, int d = 1) { return __shfl_up_sync(kFull, v, d, SEG); }

// Parallel cyclic reduction of a tridiagonal system spread over a SEG-lane segment:
// row l is (a, b, c | d) = (lower, diag, upper | rhs).  Rows outside the column are
// identity rows.  Out-of-range shuffles return the lane's own (finite) value and are
// multiplied by a coupling that is exactly zero there.
template <int SEG>
__device__ __forceinline__ double pcr_solve(double a, double b, double c, double d)
{
#pragma unroll
    for (int s = 1; s < SEG; s <<= 1) {
        const double r = 1.0 / b;
        const double al = -a * from_below<SEG>(r, s);
        const double ga = -c * from_above<SEG>(r, s);
        const double a_m = from_below<SEG>(a, s), c_m = from_below<SEG>(c, s), d_m = from_below<SEG>(d, s);
        const double a_p = from_above<SEG>(a, s), c_p = from_above<SEG>(c, s), d_p = from_above<SEG>(d, s);
        b = b + al * c_m + ga * a_p;
        d = d + al * d_m + ga * d_p;
        a = al * a_m;
        c = ga * c_p;
    }
    return d / b;
}

// The same reduction split into a matrix part (once) and a right-hand-side part (per solve).
template <int SEG>
struct PcrFactor {
    static constexpr int kSteps = (SEG == 32) ? 5 : 4;
    double al[kSteps], ga[kSteps], inv_b;
    __device__ __forceinline__ void factor(double a, double b, double c)
    {
        int k = 0;
#pragma unroll
        for (int s = 1; s < SEG; s <<= 1, ++k) {
            const double r = 1.0 / b;
            al[k] = -a * from_below<SEG>(r, s);
            ga[k] = -c * from_above<SEG>(r, s);
            const double a_m = from_below<SEG>(a, s), c_m = from_below<SEG>(c, s);
            const double a_p = from_above<SEG>(a, s), c_p = from_above<SEG>(c, s);
            b = b + al[k] * c_m + ga[k] * a_p;
            a = al[k] * a_m;
            c = ga[k] * c_p;
        }
        inv_b = 1.0 / b;
    }
    __device__ __forceinline__ double solve(double d) const
    {
        int k = 0;
#pragma unroll
        for (int s = 1; s < SEG; s <<= 1, ++k) d = d + al[k] * from_below<SEG>(d, s) + ga[k] * from_above<SEG>(d, s);
        return d * inv_b;
    }
};

// Benign parameters for pad lanes (levels >= N, columns >= ncol): finite everywhere.
__device__ __forceinline__ HydroCell dummy_cell()
{
    HydroCell h;
    h.nu = 0.5; h.theta_r = 0.1; h.K_sat = 0.0; h.S_s = 1e-3; h.a = 1.0; h.b = 2.0; h.m = 0.5;
    return h;
}

template <int CLOSURE, int MATH, int MODEL, int SEG>
__global__ void __launch_bounds__(128) k_step_warp(const DevView P, double dtg, int max_iters)
{
    constexpr int CPW = 32 / SEG;  // columns per warp
    const int lane = threadIdx.x & 31;
    const int l = lane & (SEG - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t c = warp * CPW + lane / SEG;
    const int N = P.N;
    const bool col_ok = c < P.ncol;
    const bool cell = col_ok && l < N;
    const bool is_bot = (l == 0), is_top = (l == N - 1);
    const bool interior = cell && l < N - 1;  // has a face above shared with level l+1
    const int64_t k = cell ? P.at(l, c) : 0;
    const int64_t cs = col_ok ? c : 0;
    const EarthConst &E = P.earth;

    // ---- per-lane constants of the stage ------------------------------------------------
    const HydroCell hc = cell ? load_cell(P, k) : dummy_cell();
    const double z = cell ? __ldg(P.z_c + l) : 0.0;
    const double idzc = cell ? __ldg(P.inv_dz_c + l) : 1.0;
    const double idzf_hi = interior ? __ldg(P.inv_dz_f + l + 1) : 0.0;
    double sat = 0.0, src_w = 0.0, src_e = 0.0, R_ss = 0.0, R_ess = 0.0;
    if (P.topmodel) {
        const double hg = fmax(P.h_grad[cs], kEps);
        R_ss = P.R_ss[cs];
        src_w = R_ss / hg;
        if (MODEL == 1) {
            R_ess = P.R_ess[cs];
            src_e = R_ess / hg;
        }
        sat = cell ? P.is_sat[k] : 0.0;
    }
    double top_w = P.top_bc_w[cs], bot_w = P.bot_bc_w[cs];
    const double temp1 = cell ? P.Y_theta_l[k] : 0.3;
    double U1 = temp1;
    const double tiw = P.Y_intF_w[cs];
    double Uiw = tiw;

    // EnergyHydrology: lagged fields, constant face coefficients, factored W22
    double theta_i = 0.0, rcds = 0.0, K_lag = 0.0, temp2 = 0.0, U2 = 0.0, nu_eff = hc.nu;
    double aK_hi = 0.0, aK_lo = 0.0, aC_hi = 0.0, aC_lo = 0.0, top_h = 0.0, bot_h = 0.0, tie = 0.0, Uie = 0.0;
    PcrFactor<SEG> W22;
    if (MODEL == 1) {
        theta_i = cell ? P.Y_theta_i[k] : 0.0;
        rcds = cell ? __ldg(P.rho_c_ds + k) : 1e6;
        K_lag = cell ? __ldg(P.K_lag + k) : 0.0;
        const double kap = cell ? __ldg(P.kappa_lag + k) : 0.0;
        const double tl_lag = cell ? __ldg(P.theta_l_lag + k) : 0.0;
        temp2 = cell ? P.Y_rho_e[k] : 0.0;
        U2 = temp2;
        nu_eff = hc.nu - theta_i;
        top_h = P.top_bc_h[cs];
        bot_h = P.bot_bc_h[cs];
        tie = P.Y_intF_e[cs];
        Uie = tie;
        // (shuffles are always executed by the whole warp, then selected)
        const double K_p = from_above<SEG>(K_lag), kap_p = from_above<SEG>(kap);
        aK_hi = interior ? ((K_lag + K_p) / 2.0) * idzf_hi : 0.0;
        aC_hi = interior ? ((kap + kap_p) / 2.0) * idzf_hi : 0.0;
        aK_lo = from_below<SEG>(aK_hi);
        aC_lo = from_below<SEG>(aC_hi);
        if (is_bot) {
            aK_lo = 0.0;
            aC_lo = 0.0;
        }
        // Jacobian uses the LAGGED theta_l for rho_c_s (energy_hydrology.jl:561-566)
        const double rc = 1 / volumetric_heat_capacity(tl_lag, theta_i, rcds, E);
        double rc_p = from_above<SEG>(rc);
        if (!interior) rc_p = 0.0;
        double rc_m = from_below<SEG>(rc);
        if (is_bot) rc_m = 0.0;
        double lo, di, up;
        tridiag_row(dtg, aC_lo, aC_hi, rc_m, rc, rc_p, idzc, 0.0, lo, di, up);
        if (!cell) {
            lo = 0.0; up = 0.0; di = -1.0;
        }
        W22.factor(lo, di, up);
    }

    // Richards with a MoistureStateBC top: boundary values do not depend on the iterate
    const bool bc_live = (MODEL == 0) && (P.top_bc == 1);
    double psi_bc = 0.0;
    if (bc_live) {
        double th = hc.nu;
        if (cell && is_top) th = P.theta_bc_top[cs];
        if (cell && is_bot && P.bottom_bc == 2) th = P.theta_bc_bot[cs];
        psi_bc = pressure_head<CLOSURE, MATH>(hc, th, hc.nu);  // used by lanes 0 / N-1 only
    }

    double dx2 = 0.0;
#pragma unroll 1
    for (int it = 0; it < max_iters; ++it) {
        // ---- cache_imp!: closures at the iterate --------------------------------------------
        double K = K_lag, psi, dps;
        if (MODEL == 0)
            closure_eval<CLOSURE, MATH, true, true, true>(hc, U1, hc.nu, K, psi, dps);
        else
            closure_eval<CLOSURE, MATH, false, true, true>(hc, U1, nu_eff, K, psi, dps);
        const double h = psi + z;
        const double h_p = from_above<SEG>(h);
        double dps_p = from_above<SEG>(dps);
        if (!interior) dps_p = 0.0;
        double dps_m = from_below<SEG>(dps);
        if (is_bot) dps_m = 0.0;
        double top_dflux = 0.0;
        if (MODEL == 0) {
            const double K_p = from_above<SEG>(K);
            aK_hi = interior ? ((K + K_p) / 2.0) * idzf_hi : 0.0;
            aK_lo = from_below<SEG>(aK_hi);
            if (is_bot) aK_lo = 0.0;
            if (bc_live) {
                if (is_top) {
                    top_w = -K * ((psi_bc + P.dz_top) - psi) / P.dz_top;
                    top_dflux = K * dps / P.dz_top;
                }
                if (is_bot) {
                    if (P.bottom_bc == 1)
                        bot_w = -1 * K;
                    else if (P.bottom_bc == 2)
                        bot_w = -K * ((psi + P.dz_bot) - psi_bc) / P.dz_bot;
                }
                top_w = __shfl_sync(kFull, top_w, N - 1, SEG);
                bot_w = __shfl_sync(kFull, bot_w, 0, SEG);
            }
        }
        // ---- T_imp!: face fluxes (owned by the lane below the face) ---------------------------
        const double dh = h_p - h;
        const double qw_hi = interior ? -aK_hi * dh : top_w;
        double qw_lo = from_below<SEG>(qw_hi);
        if (is_bot) qw_lo = bot_w;
        double Tw = -((qw_hi - qw_lo) * idzc);
        if (P.topmodel) Tw -= src_w * sat;
        double f1 = temp1 + dtg * Tw - U1;
        // ---- Wfact: (theta_l, theta_l) -------------------------------------------------------
        double lo, di, up;
        tridiag_row(dtg, aK_lo, aK_hi, dps_m, dps, dps_p, idzc, top_dflux, lo, di, up);
        if (!cell) {
            lo = 0.0; up = 0.0; di = -1.0; f1 = 0.0;
        }
        double f2 = 0.0, aE_hi = 0.0, aE_lo = 0.0;
        if (MODEL == 1) {
            const double T = eh_temperature(U1, U2, theta_i, hc.nu, rcds, E);
            const double eK = volumetric_internal_energy_liq(T, E) * K_lag;
            const double eK_p = from_above<SEG>(eK);
            aE_hi = interior ? ((eK + eK_p) / 2.0) * idzf_hi : 0.0;
            aE_lo = from_below<SEG>(aE_hi);
            if (is_bot) aE_lo = 0.0;
            const double T_p = from_above<SEG>(T);
            const double qe_hi = interior ? -aC_hi * (T_p - T) - aE_hi * dh : top_h;
            double qe_lo = from_below<SEG>(qe_hi);
            if (is_bot) qe_lo = bot_h;
            double Te = -((qe_hi - qe_lo) * idzc);
            if (P.topmodel) Te -= src_e * sat;
            f2 = cell ? temp2 + dtg * Te - U2 : 0.0;
        }
        // ---- ldiv!: BlockDiagonalSolve / BlockLowerTriangularSolve(theta_l) --------------------
        const double x1 = pcr_solve<SEG>(lo, di, up, f1);
        U1 -= x1;
        dx2 = cell ? x1 * x1 : 0.0;
        if (MODEL == 1) {
            const double y = cell ? dps * x1 : 0.0;
            const double y_p = from_above<SEG>(y), y_m = from_below<SEG>(y);
            // (W21 x1) with W21 = -dtg*(D . Diag(interp(-eK)) . G . Diag(dpsi)) - I  (energy_hydrology.jl:545-556)
            const double s = dtg * ((aE_lo * (y_m - y) + aE_hi * (y_p - y)) * idzc) - x1;
            const double x2 = W22.solve(cell ? f2 - s : 0.0);
            U2 -= x2;
            if (cell) dx2 += x2 * x2;
        }
        // ---- flux integrals (W = -I): lane 0 of the segment keeps them -------------------------
        double Tiw = -(top_w - bot_w);
        if (P.topmodel) Tiw -= R_ss;
        const double dxw = -(tiw + dtg * Tiw - Uiw);
        Uiw -= dxw;
        if (is_bot && col_ok) dx2 += dxw * dxw;
        if (MODEL == 1) {
            double Tie = -(top_h - bot_h);
            if (P.topmodel) Tie -= R_ess;
            const double dxe = -(tie + dtg * Tie - Uie);
            Uie -= dxe;
            if (is_bot && col_ok) dx2 += dxe * dxe;
        }
    }

    // ---- write the new state ---------------------------------------------------------------
    double bad = 0.0;
    if (cell) {
        P.out_theta_l[k] = U1;
        if (!isfinite(U1)) bad += 1.0;
        if (MODEL == 1) {
            P.out_rho_e[k] = U2;
            if (!isfinite(U2)) bad += 1.0;
        }
        if (is_bot) {
            P.out_intF_w[c] = Uiw;
            if (MODEL == 1) P.out_intF_e[c] = Uie;
            if (bc_live) {
                P.top_bc_w[c] = top_w;
                P.bot_bc_w[c] = bot_w;
            }
        }
    }
    accumulate_stats(P, dx2, bad);
}

}  // namespace clb
