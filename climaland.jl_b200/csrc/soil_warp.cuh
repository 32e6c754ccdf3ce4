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

#ifndef CLB_WARP_MIN_BLOCKS
#define CLB_WARP_MIN_BLOCKS 4
#endif
constexpr unsigned kFull = 0xffffffffu;

template <int SEG>
__device__ __forceinline__ double from_above(double v, int d = 1) { return __shfl_down_sync(kFull, v, d, SEG); }
template <int SEG>
__device__ __forceinline__ double from_below(double v, int d = 1) { return __shfl_up_sync(kFull, v, d, SEG); }

// Parallel cyclic reduction of a tridiagonal system spread over a SEG-lane segment:
// row l is (a, b, c | d) = (lower, diag, upper | rhs).  Rows outside the column are
// identity rows.  Out-of-range shuffles return the lane's own (finite) value and are
// multiplied by a coupling that is exactly zero there.
// Rows are kept normalised (diagonal = 1), so a step needs three values from each
// neighbour (a, c, d) and one reciprocal.
template <int SEG>
__device__ __forceinline__ double pcr_solve(double a, double b, double c, double d)
{
    double r = fm::rcp(b);
    a *= r; c *= r; d *= r;
#pragma unroll
    for (int s = 1; s < SEG; s <<= 1) {
        const double a_m = from_below<SEG>(a, s), c_m = from_below<SEG>(c, s), d_m = from_below<SEG>(d, s);
        const double a_p = from_above<SEG>(a, s), c_p = from_above<SEG>(c, s), d_p = from_above<SEG>(d, s);
        r = fm::rcp(fma(-a, c_m, fma(-c, a_p, 1.0)));
        d = fma(-a, d_m, fma(-c, d_p, d)) * r;
        a = -(a * a_m) * r;
        c = -(c * c_p) * r;
    }
    return d;
}

// The same reduction split into a matrix part (once) and a right-hand-side part (per solve).
template <int SEG>
struct PcrFactor {
    static constexpr int kSteps = (SEG == 32) ? 5 : 4;
    double a_s[kSteps], c_s[kSteps], r_s[kSteps + 1];
    __device__ __forceinline__ void factor(double a, double b, double c)
    {
        double r = fm::rcp(b);
        r_s[0] = r;
        a *= r; c *= r;
        int k = 0;
#pragma unroll
        for (int s = 1; s < SEG; s <<= 1, ++k) {
            const double a_m = from_below<SEG>(a, s), c_m = from_below<SEG>(c, s);
            const double a_p = from_above<SEG>(a, s), c_p = from_above<SEG>(c, s);
            a_s[k] = a;
            c_s[k] = c;
            r = fm::rcp(fma(-a, c_m, fma(-c, a_p, 1.0)));
            r_s[k + 1] = r;
            a = -(a * a_m) * r;
            c = -(c * c_p) * r;
        }
    }
    __device__ __forceinline__ double solve(double d) const
    {
        d *= r_s[0];
        int k = 0;
#pragma unroll
        for (int s = 1; s < SEG; s <<= 1, ++k)
            d = fma(-a_s[k], from_below<SEG>(d, s), fma(-c_s[k], from_above<SEG>(d, s), d)) * r_s[k + 1];
        return d;
    }
};

// Benign parameters for pad lanes (levels >= N, columns >= ncol): finite everywhere.
__device__ __forceinline__ HydroCell dummy_cell()
{
    HydroCell h;
    h.nu = 0.5; h.theta_r = 0.1; h.K_sat = 0.0; h.S_s = 1e-3; h.a = 1.0; h.b = 2.0; h.m = 0.5;
    return h;
}

// IEEE operations in LIBM mode, the branch-free ones of soil_math.cuh in FAST mode
template <int MATH>
__device__ __forceinline__ double m_rcp(double x) { return MATH == kMathLibm ? 1.0 / x : fm::rcp(x); }
template <int MATH>
__device__ __forceinline__ double m_div(double a, double b) { return MATH == kMathLibm ? a / b : fm::div(a, b); }

// Neighbour exchange by rotation within the segment: the last lane of a segment (SEG - 1)
// is a GHOST lane -- the host dispatches SEG > N -- that holds the column's bottom boundary
// values, so lane 0's "below" neighbour is the ghost and no boundary select is needed.
template <int SEG>
__device__ __forceinline__ double rot_above(double v, int l) { return __shfl_sync(kFull, v, (l + 1) & (SEG - 1), SEG); }
template <int SEG>
__device__ __forceinline__ double rot_below(double v, int l) { return __shfl_sync(kFull, v, (l - 1) & (SEG - 1), SEG); }

template <int CLOSURE, int MATH, int MODEL, int SEG>
__global__ void __launch_bounds__(128, CLB_WARP_MIN_BLOCKS) k_step_warp(const DevView P, double dtg, int max_iters)
{
    constexpr int CPW = 32 / SEG;  // columns per warp
    constexpr int G = SEG - 1;     // ghost lane
    const int lane = threadIdx.x & 31;
    const int l = lane & (SEG - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t c = warp * CPW + lane / SEG;
    const int N = P.N;  // N <= SEG - 1
    const bool col_ok = c < P.ncol;
    const bool cell = col_ok && l < N;
    const bool is_bot = (l == 0), is_top = (l == N - 1), is_ghost = (l == G);
    const bool interior = cell && l < N - 1;  // has a face above shared with level l+1
    const EarthConst &E = P.earth;
    // pad lanes read a valid cell (clamped indices) and are neutralised below
    const int ls = min(l, N - 1);
    const int64_t cs = col_ok ? c : P.ncol - 1;
    const int64_t k = P.at(ls, cs);

    // ---- every global load of the stage, issued back to back (one DRAM round trip);
    //      fields the configuration does not use point at zero-filled dummies ----------------
    HydroCell hc = load_cell(P, k);
    const double z = __ldg(P.z_c + ls);
    double idzc = __ldg(P.inv_dz_c + ls);
    double idzf_hi = __ldg(P.inv_dz_f + min(ls + 1, N - 1));
    const double ld_theta = P.Y_theta_l[k];
    const double ld_sat = P.is_sat[k];
    const double ld_Rss = P.R_ss[cs];
    const double ld_hg = P.h_grad[cs];
    double top_w = P.top_bc_w[cs], bot_w = P.bot_bc_w[cs];
    double ld_theta_i = 0.0, ld_rcds = 1e6, ld_K = 0.0, ld_kap = 0.0, ld_tl = 0.0, ld_rho_e = 0.0, ld_Ress = 0.0;
    double top_h = 0.0, bot_h = 0.0;
    if (MODEL == 1) {
        ld_theta_i = P.Y_theta_i[k];
        ld_rcds = __ldg(P.rho_c_ds + k);
        ld_K = __ldg(P.K_lag + k);
        ld_kap = __ldg(P.kappa_lag + k);
        ld_tl = __ldg(P.theta_l_lag + k);
        ld_rho_e = P.Y_rho_e[k];
        ld_Ress = P.R_ess[cs];
        top_h = P.top_bc_h[cs];
        bot_h = P.bot_bc_h[cs];
    }
    const bool bc_live = (MODEL == 0) && (P.top_bc == 1);  // cache holds dfluxBCdY (rre.jl:460-468)
    double ld_thbc = hc.nu;
    if (bc_live) {
        if (is_top) ld_thbc = P.theta_bc_top[cs];
        if (is_bot && P.bottom_bc == 2) ld_thbc = P.theta_bc_bot[cs];
    }
    const double tiw = P.Y_intF_w[cs];
    double Uiw = tiw;
    const double tie = (MODEL == 1) ? P.Y_intF_e[cs] : 0.0;

    // ---- per-lane constants of the stage.  Pad and ghost lanes get idzc = 0, no source and a
    //      zero face coefficient: their rows are identity rows and their update is exactly 0 ----
    if (!cell) {
        hc = dummy_cell();
        idzc = 0.0;
    }
    if (!interior) idzf_hi = 0.0;
    const double temp1 = cell ? ld_theta : 0.3;
    double U1 = temp1;
    // implicit TOPMODEL source (Runoff/Runoff.jl:321-359): per-cell constants src * is_saturated
    const double inv_hg = m_rcp<MATH>(fmax(ld_hg, kEps));
    const double sw = cell ? (ld_Rss * inv_hg) * ld_sat : 0.0;
    const double se = (MODEL == 1 && cell) ? (ld_Ress * inv_hg) * ld_sat : 0.0;
    // boundary face fluxes enter as an additive constant of the face above: the top lane owns
    // the top boundary face, the ghost lane "owns" the bottom boundary face
    double qadd_w = 0.0, qadd_e = 0.0;
    if (col_ok && is_top) { qadd_w = top_w; qadd_e = top_h; }
    if (col_ok && is_ghost) { qadd_w = bot_w; qadd_e = bot_h; }

    // EnergyHydrology: lagged fields, constant face coefficients, factored W22
    const double theta_i = (MODEL == 1 && cell) ? ld_theta_i : 0.0;
    const double rcds = ld_rcds;
    const double K_lag = cell ? ld_K : 0.0;
    const double temp2 = cell ? ld_rho_e : 0.0;
    double U2 = temp2;
    const double nu_eff = hc.nu - theta_i;  // nu for Richards
    double aK_hi = 0.0, aK_lo = 0.0, aC_hi = 0.0;
    PcrFactor<SEG> W22;
    if (MODEL == 1) {
        const double kap = cell ? ld_kap : 0.0;
        aK_hi = ((K_lag + rot_above<SEG>(K_lag, l)) / 2.0) * idzf_hi;
        aC_hi = ((kap + rot_above<SEG>(kap, l)) / 2.0) * idzf_hi;
        aK_lo = rot_below<SEG>(aK_hi, l);
        const double aC_lo = rot_below<SEG>(aC_hi, l);
        // Jacobian uses the LAGGED theta_l for rho_c_s (energy_hydrology.jl:561-566)
        const double rc = m_rcp<MATH>(volumetric_heat_capacity(ld_tl, theta_i, rcds, E));
        const double rc_p = rot_above<SEG>(rc, l), rc_m = rot_below<SEG>(rc, l);
        double lo, di, up;
        tridiag_row(dtg, aC_lo, aC_hi, rc_m, rc, rc_p, idzc, 0.0, lo, di, up);
        W22.factor(lo, di, up);
    }
    double dx2_int = 0.0;
    if (MODEL == 1) {
        // flux integrals (W = -I) with lagged boundary fluxes do not depend on the iterate: run the
        // same Newton recurrence here, off the hot loop, so its operands die before it
        const double Tiw = -(top_w - bot_w) - ld_Rss, Tie = -(top_h - bot_h) - ld_Ress;
        double Ue = tie, dxw = 0.0, dxe = 0.0;
        for (int it = 0; it < max_iters; ++it) {
            dxw = -(tiw + dtg * Tiw - Uiw);
            Uiw -= dxw;
            dxe = -(tie + dtg * Tie - Ue);
            Ue -= dxe;
        }
        if (is_bot && col_ok) {
            dx2_int = dxw * dxw + dxe * dxe;
            P.out_intF_w[c] = Uiw;
            P.out_intF_e[c] = Ue;
        }
    }

    // Richards with a MoistureStateBC top: boundary heads do not depend on the iterate
    double psi_bc = 0.0;
    if (bc_live) psi_bc = pressure_head<CLOSURE, MATH>(hc, ld_thbc, hc.nu);  // used by lanes 0 / N-1 only

    const CellEval<CLOSURE, MATH> ce(hc, nu_eff);
    double dx2 = 0.0;
#pragma unroll 1
    for (int it = 0; it < max_iters; ++it) {
        // ---- cache_imp!: closures at the iterate --------------------------------------------
        double K = K_lag, psi, dps;
        if (MODEL == 0)
            ce.template eval<true, true, true>(U1, K, psi, dps);
        else
            ce.template eval<false, true, true>(U1, K, psi, dps);
        const double h = psi + z;
        const double dh = rot_above<SEG>(h, l) - h;
        const double dps_p = rot_above<SEG>(dps, l), dps_m = rot_below<SEG>(dps, l);
        double top_dflux = 0.0;
        if (MODEL == 0) {
            aK_hi = ((K + rot_above<SEG>(K, l)) / 2.0) * idzf_hi;
            aK_lo = rot_below<SEG>(aK_hi, l);
            if (bc_live) {
                double qb = 0.0;  // bottom boundary flux, evaluated by lane 0, owned by the ghost
                if (P.bottom_bc == 1)
                    qb = -1 * K;
                else if (P.bottom_bc == 2)
                    qb = fm::div_by(-K * ((psi + P.dz_bot) - psi_bc), P.dz_bot, P.inv_dz_bot);
                else
                    qb = bot_w;
                qb = rot_above<SEG>(qb, l);
                if (col_ok && is_ghost) qadd_w = qb;
                if (col_ok && is_top) {
                    qadd_w = fm::div_by(-K * ((psi_bc + P.dz_top) - psi), P.dz_top, P.inv_dz_top);
                    top_dflux = fm::div_by(K * dps, P.dz_top, P.inv_dz_top);
                }
                top_w = __shfl_sync(kFull, qadd_w, N - 1, SEG);
                bot_w = __shfl_sync(kFull, qadd_w, G, SEG);
            }
        }
        // ---- T_imp!: face fluxes (owned by the lane below the face) ---------------------------
        const double qw_hi = fma(-aK_hi, dh, qadd_w);
        const double qw_lo = rot_below<SEG>(qw_hi, l);
        const double Tw = -((qw_hi - qw_lo) * idzc) - sw;
        const double f1 = temp1 + dtg * Tw - U1;
        // ---- Wfact: (theta_l, theta_l) -------------------------------------------------------
        double lo, di, up;
        tridiag_row(dtg, aK_lo, aK_hi, dps_m, dps, dps_p, idzc, top_dflux, lo, di, up);
        double f2 = 0.0, aE_hi = 0.0, aE_lo = 0.0;
        if (MODEL == 1) {
            const double T = eh_temperature_m<MATH>(U1, U2, theta_i, hc.nu, rcds, E);
            const double eK = volumetric_internal_energy_liq(T, E) * K_lag;
            aE_hi = ((eK + rot_above<SEG>(eK, l)) / 2.0) * idzf_hi;
            aE_lo = rot_below<SEG>(aE_hi, l);
            const double dT = rot_above<SEG>(T, l) - T;
            const double qe_hi = fma(-aC_hi, dT, fma(-aE_hi, dh, qadd_e));
            const double qe_lo = rot_below<SEG>(qe_hi, l);
            const double Te = -((qe_hi - qe_lo) * idzc) - se;
            f2 = temp2 + dtg * Te - U2;
        }
        // ---- ldiv!: BlockDiagonalSolve / BlockLowerTriangularSolve(theta_l) --------------------
        const double x1 = pcr_solve<SEG>(lo, di, up, f1);
        U1 -= x1;
        dx2 = x1 * x1;
        if (MODEL == 1) {
            const double y = dps * x1;
            const double y_p = rot_above<SEG>(y, l), y_m = rot_below<SEG>(y, l);
            // (W21 x1) with W21 = -dtg*(D . Diag(interp(-eK)) . G . Diag(dpsi)) - I  (energy_hydrology.jl:545-556)
            const double s = dtg * ((aE_lo * (y_m - y) + aE_hi * (y_p - y)) * idzc) - x1;
            const double x2 = W22.solve(f2 - s);
            U2 -= x2;
            dx2 += x2 * x2;
        }
        if (MODEL == 0) {
            // flux integral (W = -I); the boundary fluxes may follow the iterate (MoistureStateBC)
            const double Tiw = -(top_w - bot_w) - ld_Rss;
            const double dxw = -(tiw + dtg * Tiw - Uiw);
            Uiw -= dxw;
            if (is_bot && col_ok) dx2 += dxw * dxw;
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
        if (is_bot && MODEL == 0) {
            P.out_intF_w[c] = Uiw;
            if (bc_live) {
                P.top_bc_w[c] = top_w;
                P.bot_bc_w[c] = bot_w;
            }
        }
    }
    dx2 += dx2_int;
    accumulate_stats(P, dx2, bad);
}

}  // namespace clb
