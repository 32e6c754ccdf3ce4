// soil_fused.cuh -- the fused implicit ARS111 stage: cache_imp! + max_iters x
// (Wfact, T_imp!, residual, ldiv!, update) in ONE kernel, one thread per column.
//
// Reference call sequence (ClimaTimeSteppers NewtonsMethod configured at
// src/simulations/Simulations.jl:127-135):
//     temp = U ; cache_imp!(U)
//     for n in 1:max_iters
//         Wfact(W, U, dtgamma) ; f = temp + dtgamma*T_imp(U) - U ; dx = W \ f ; U -= dx
//         n < max_iters && cache_imp!(U)
// Per column nothing but the state leaves the SM: K, psi, dpsi/dtheta, T, the
// tridiagonal rows and the Thomas work vectors live in registers (static N) or in a
// column-fastest scratch (runtime N).  Levels are swept bottom -> top with a
// one-level lookahead, so each closure is evaluated exactly once per Newton iteration.
#pragma once
#include "soil_device.cuh"
#include "soil_hooks.cuh"

namespace clb {

// ---- column storage policies ------------------------------------------------
template <int NS>
struct RegCol {  // registers; requires fully unrolled level loops
    double v[NS];
    __device__ __forceinline__ double get(int i) const { return v[i]; }
    __device__ __forceinline__ void set(int i, double x) { v[i] = x; }
};
struct MemCol {  // global scratch with the mirrors' level stride
    double *p;
    int64_t ld;
    __device__ __forceinline__ double get(int i) const { return p[(int64_t)i * ld]; }
    __device__ __forceinline__ void set(int i, double x) { p[(int64_t)i * ld] = x; }
};

// grid access: compile-time (kernel-parameter constants) or runtime (device arrays)
template <int NS>
struct GridS {
    const GridConst<NS> &g;
    __device__ __forceinline__ double z(int i) const { return g.z_c[i]; }
    __device__ __forceinline__ double idzc(int i) const { return g.inv_dz_c[i]; }
    __device__ __forceinline__ double idzf(int i) const { return g.inv_dz_f[i]; }
};
struct GridR {
    const double *zc, *ic, *jf;
    __device__ __forceinline__ double z(int i) const { return __ldg(zc + i); }
    __device__ __forceinline__ double idzc(int i) const { return __ldg(ic + i); }
    __device__ __forceinline__ double idzf(int i) const { return __ldg(jf + i); }
};

// Warp-then-atomic accumulation of the step statistics.
__device__ __forceinline__ void accumulate_stats(const DevView &P, double dx2, double bad)
{
    for (int o = 16; o > 0; o >>= 1) {
        dx2 += __shfl_xor_sync(0xffffffffu, dx2, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(P.stats + 0, dx2);
        if (bad != 0.0) atomicAdd(P.stats + 1, bad);
    }
}

// ---- Richards ---------------------------------------------------------------
// One Newton iteration of appendix A.1 for one column.  U holds the iterate; `temp`
// is read from the untouched state field.  Returns sum(dx^2) of this iteration.
template <int CLOSURE, int MATH, int NS, class Col, class Grid>
__device__ __forceinline__ double richards_newton_iteration(const DevView &P, const Grid &G, int64_t c, double dtg,
                                                            Col &U, Col &cp, Col &dp, bool first, double psi_bc_top,
                                                            double psi_bc_bot, double &top_w, double &bot_w,
                                                            double &Uint, double temp_int)
{
    const int N = NS > 0 ? NS : P.N;
    const bool bc_live = (P.top_bc == 1);  // cache holds dfluxBCdY -> BCs re-evaluated (rre.jl:460-468)
    double src_w = 0.0, R_ss = 0.0;
    if (P.topmodel) {
        R_ss = P.R_ss[c];
        src_w = R_ss / fmax(P.h_grad[c], kEps);
    }
    // level 0
    HydroCell cell = load_cell(P, P.at(0, c));
    double K0, psi0, d0;
    CellEval<CLOSURE, MATH>(cell, cell.nu).template eval<true, true, true>(U.get(0), K0, psi0, d0);
    if (bc_live) {
        if (P.bottom_bc == 1)
            bot_w = -1 * K0;
        else if (P.bottom_bc == 2)
            bot_w = -K0 * ((psi0 + P.dz_bot) - psi_bc_bot) / P.dz_bot;
    }
    double a_lo = 0.0, q_lo = bot_w, d_m = 0.0;
    double cprev = 0.0, dprev = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int64_t k = P.at(i, c);
        double a_hi = 0.0, q_hi, K1 = 0.0, psi1 = 0.0, d1 = 0.0, top_dflux = 0.0;
        if (i < N - 1) {
            cell = load_cell(P, k + P.sl);
            CellEval<CLOSURE, MATH>(cell, cell.nu).template eval<true, true, true>(U.get(i + 1), K1, psi1, d1);
            a_hi = ((K0 + K1) / 2.0) * G.idzf(i + 1);
            q_hi = -a_hi * ((psi1 + G.z(i + 1)) - (psi0 + G.z(i)));
        } else {
            if (bc_live) {
                top_w = -K0 * ((psi_bc_top + P.dz_top) - psi0) / P.dz_top;
                top_dflux = K0 * d0 / P.dz_top;
            }
            q_hi = top_w;
        }
        const double idzc = G.idzc(i);
        double lo, di, up;
        tridiag_row(dtg, a_lo, a_hi, d_m, d0, d1, idzc, top_dflux, lo, di, up);
        double T = -((q_hi - q_lo) * idzc);
        if (P.topmodel) T -= src_w * P.is_sat[k];
        const double u = U.get(i);
        const double tmp = first ? u : P.Y_theta_l[k];
        const double f = tmp + dtg * T - u;
        // Thomas forward elimination
        const double den = 1.0 / (di - lo * cprev);
        cprev = up * den;
        dprev = (f - lo * dprev) * den;
        cp.set(i, cprev);
        dp.set(i, dprev);
        a_lo = a_hi; q_lo = q_hi; d_m = d0;
        K0 = K1; psi0 = psi1; d0 = d1;
    }
    // back substitution and update
    double x = dprev, dx2 = x * x;
    U.set(N - 1, U.get(N - 1) - x);
#pragma unroll
    for (int i = N - 2; i >= 0; --i) {
        x = dp.get(i) - cp.get(i) * x;
        dx2 += x * x;
        U.set(i, U.get(i) - x);
    }
    // flux integral: W = -I, so dx = -f  (rre.jl:166; implicit_timestepping.jl:149-152)
    double Tint = -(top_w - bot_w);
    if (P.topmodel) Tint -= R_ss;
    const double fint = temp_int + dtg * Tint - Uint;
    const double dxint = -fint;
    Uint -= dxint;
    return dx2 + dxint * dxint;
}

template <int CLOSURE, int MATH>
__device__ __forceinline__ void richards_bc_constants(const DevView &P, int64_t c, double &psi_bc_top,
                                                      double &psi_bc_bot)
{
    psi_bc_top = 0.0;
    psi_bc_bot = 0.0;
    if (P.top_bc == 1) {  // state BC values do not depend on the iterate: once per stage
        const HydroCell ct = load_cell(P, P.at(P.N - 1, c));
        psi_bc_top = pressure_head<CLOSURE, MATH>(ct, P.theta_bc_top[c], ct.nu);
        if (P.bottom_bc == 2) {
            const HydroCell cb = load_cell(P, P.at(0, c));
            psi_bc_bot = pressure_head<CLOSURE, MATH>(cb, P.theta_bc_bot[c], cb.nu);
        }
    }
}

// Register-column variant: static N, the iterate and the Thomas vectors in registers.
template <int CLOSURE, int MATH, int NS>
__global__ void __launch_bounds__(128) k_richards_step_reg(const DevView P, const GridConst<NS> gc, double dtg,
                                                           int max_iters)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < P.ncol;
    double dx2 = 0.0, bad = 0.0;
    if (live) {
        const GridS<NS> G{gc};
        RegCol<NS> U, cp, dp;
#pragma unroll
        for (int i = 0; i < NS; ++i) U.set(i, P.Y_theta_l[P.at(i, c)]);
        double top_w = P.top_bc_w[c], bot_w = P.bot_bc_w[c];
        const double temp_int = P.Y_intF_w[c];
        double Uint = temp_int;
        double psi_bc_top, psi_bc_bot;
        richards_bc_constants<CLOSURE, MATH>(P, c, psi_bc_top, psi_bc_bot);
#pragma unroll 1
        for (int it = 0; it < max_iters; ++it)
            dx2 = richards_newton_iteration<CLOSURE, MATH, NS>(P, G, c, dtg, U, cp, dp, it == 0, psi_bc_top,
                                                               psi_bc_bot, top_w, bot_w, Uint, temp_int);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double u = U.get(i);
            P.out_theta_l[P.at(i, c)] = u;
            if (!isfinite(u)) bad += 1.0;
        }
        P.out_intF_w[c] = Uint;
        if (P.top_bc == 1) {  // the cache keeps the last evaluated boundary fluxes
            P.top_bc_w[c] = top_w;
            P.bot_bc_w[c] = bot_w;
        }
    }
    accumulate_stats(P, dx2, bad);
}

// Generic variant: runtime N, iterate and Thomas vectors in a column-fastest scratch.
// iter_begin/iter_end select the Newton iterations this launch performs: the
// fixed-iteration path runs [0, max_iters) in one launch; the tolerance path launches
// one iteration at a time and skips the work once *converged is set.
template <int CLOSURE, int MATH>
__global__ void __launch_bounds__(128) k_richards_step_generic(const DevView P, double dtg, int iter_begin,
                                                               int iter_end)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < P.ncol && !(P.converged && *P.converged);
    double dx2 = 0.0, bad = 0.0;
    if (live) {
        const GridR G{P.z_c, P.inv_dz_c, P.inv_dz_f};
        const int N = P.N;
        const int64_t o = P.at(0, c);
        MemCol U{P.work[0] + o, P.sl}, cp{P.work[1] + o, P.sl}, dp{P.work[2] + o, P.sl};
        double *carry = P.carry + c;  // per-column scalars carried between launches
        double top_w, bot_w, Uint;
        const double temp_int = P.Y_intF_w[c];
        if (iter_begin == 0) {
            for (int i = 0; i < N; ++i) U.set(i, P.Y_theta_l[P.at(i, c)]);
            top_w = P.top_bc_w[c];
            bot_w = P.bot_bc_w[c];
            Uint = temp_int;
        } else {
            top_w = carry[0];
            bot_w = carry[P.ld];
            Uint = carry[2 * P.ld];
        }
        double psi_bc_top, psi_bc_bot;
        richards_bc_constants<CLOSURE, MATH>(P, c, psi_bc_top, psi_bc_bot);
        for (int it = iter_begin; it < iter_end; ++it)
            dx2 = richards_newton_iteration<CLOSURE, MATH, 0>(P, G, c, dtg, U, cp, dp, it == 0, psi_bc_top,
                                                              psi_bc_bot, top_w, bot_w, Uint, temp_int);
        carry[0] = top_w;
        carry[P.ld] = bot_w;
        carry[2 * P.ld] = Uint;
    }
    accumulate_stats(P, dx2, bad);
}

// Commit of the generic / tolerance path: scratch iterate -> state, NaN count.
__global__ void __launch_bounds__(128) k_commit_state(const DevView P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double bad = 0.0;
    if (c < P.ncol) {
        const bool eh = (P.model == 1);
        for (int i = 0; i < P.N; ++i) {
            const int64_t k = P.at(i, c);
            const double u = P.work[0][k];
            P.out_theta_l[k] = u;
            if (!isfinite(u)) bad += 1.0;
            if (eh) {
                const double e = P.work[1][k];
                P.out_rho_e[k] = e;
                if (!isfinite(e)) bad += 1.0;
            }
        }
        const double *carry = P.carry + c;
        if (!eh) {
            P.out_intF_w[c] = carry[2 * P.ld];
            if (P.top_bc == 1) {
                P.top_bc_w[c] = carry[0];
                P.bot_bc_w[c] = carry[P.ld];
            }
        } else {
            P.out_intF_w[c] = carry[0];
            P.out_intF_e[c] = carry[P.ld];
        }
    }
    accumulate_stats(P, 0.0, bad);
}

// ---- EnergyHydrology ----------------------------------------------------------
// One Newton iteration of appendix A.2 for one column: K, kappa, theta_l (for the
// Jacobian's rho_c_s) and the boundary fluxes are lagged inputs; psi, dpsi, T follow the
// iterate.  Solve = BlockLowerTriangularSolve(theta_l): Thomas(W11), mat-vec with W21
// (including its "- I", energy_hydrology.jl:554-556), Thomas(W22).
template <int CLOSURE, int MATH, int NS, class Col, class Grid>
__device__ __forceinline__ double eh_newton_iteration(const DevView &P, const Grid &G, int64_t c, double dtg, Col &U1,
                                                      Col &U2, Col &cp, Col &dp, Col &D, Col &F2, bool first,
                                                      double &Uintw, double &Uinte, double temp_intw,
                                                      double temp_inte)
{
    const int N = NS > 0 ? NS : P.N;
    const EarthConst &E = P.earth;
    const double top_w = P.top_bc_w[c], bot_w = P.bot_bc_w[c];
    const double top_h = P.top_bc_h[c], bot_h = P.bot_bc_h[c];
    double src_w = 0.0, src_e = 0.0, R_ss = 0.0, R_ess = 0.0;
    if (P.topmodel) {
        const double hg = fmax(P.h_grad[c], kEps);
        R_ss = P.R_ss[c];
        R_ess = P.R_ess[c];
        src_w = R_ss / hg;
        src_e = R_ess / hg;
    }
    // ---- sweep 1: water rows + residuals, Thomas forward on W11; energy residual stored
    auto level = [&](int i, double &psi, double &dps, double &T, double &K, double &kap) {
        const int64_t k = P.at(i, c);
        const HydroCell cell = load_cell(P, k);
        const double theta_i = P.Y_theta_i[k];
        const double theta = U1.get(i);
        double Kdummy;
        CellEval<CLOSURE, MATH>(cell, cell.nu - theta_i).template eval<false, true, true>(theta, Kdummy, psi, dps);
        T = eh_temperature(theta, U2.get(i), theta_i, cell.nu, __ldg(P.rho_c_ds + k), E);
        K = __ldg(P.K_lag + k);
        kap = __ldg(P.kappa_lag + k);
    };
    double psi0, d0, T0, K0, kap0;
    level(0, psi0, d0, T0, K0, kap0);
    double eK0 = volumetric_internal_energy_liq(T0, E) * K0;
    double aK_lo = 0.0, d_m = 0.0, qw_lo = bot_w, qe_lo = bot_h;
    double cprev = 0.0, dprev = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int64_t k = P.at(i, c);
        double psi1 = 0, d1 = 0, T1 = 0, K1 = 0, kap1 = 0, eK1 = 0;
        double aK_hi = 0.0, qw_hi, qe_hi;
        if (i < N - 1) {
            level(i + 1, psi1, d1, T1, K1, kap1);
            eK1 = volumetric_internal_energy_liq(T1, E) * K1;
            const double idzf = G.idzf(i + 1);
            aK_hi = ((K0 + K1) / 2.0) * idzf;
            const double dh = (psi1 + G.z(i + 1)) - (psi0 + G.z(i));
            qw_hi = -aK_hi * dh;
            qe_hi = -(((kap0 + kap1) / 2.0) * idzf) * (T1 - T0) - (((eK0 + eK1) / 2.0) * idzf) * dh;
        } else {
            qw_hi = top_w;
            qe_hi = top_h;
        }
        const double idzc = G.idzc(i);
        double lo, di, up;
        tridiag_row(dtg, aK_lo, aK_hi, d_m, d0, d1, idzc, 0.0, lo, di, up);
        double Tw = -((qw_hi - qw_lo) * idzc);
        double Te = -((qe_hi - qe_lo) * idzc);
        if (P.topmodel) {
            const double sat = P.is_sat[k];
            Tw -= src_w * sat;
            Te -= src_e * sat;
        }
        const double u1 = U1.get(i), u2 = U2.get(i);
        const double t1 = first ? u1 : P.Y_theta_l[k];
        const double t2 = first ? u2 : P.Y_rho_e[k];
        const double f1 = t1 + dtg * Tw - u1;
        F2.set(i, t2 + dtg * Te - u2);
        D.set(i, d0);
        const double den = 1.0 / (di - lo * cprev);
        cprev = up * den;
        dprev = (f1 - lo * dprev) * den;
        cp.set(i, cprev);
        dp.set(i, dprev);
        aK_lo = aK_hi; d_m = d0; qw_lo = qw_hi; qe_lo = qe_hi;
        psi0 = psi1; d0 = d1; T0 = T1; K0 = K1; kap0 = kap1; eK0 = eK1;
    }
    // ---- back substitution 1: x1 -> dp, y = dpsi*x1 -> D
    double x = dprev, dx2 = x * x;
    D.set(N - 1, D.get(N - 1) * x);
#pragma unroll
    for (int i = N - 2; i >= 0; --i) {
        x = dp.get(i) - cp.get(i) * x;
        dp.set(i, x);
        D.set(i, D.get(i) * x);
        dx2 += x * x;
    }
    // ---- sweep 2: b2' = f2 - W21 x1, rows of W22, Thomas forward (c' -> cp, d' -> F2)
    auto level2 = [&](int i, double &eK, double &kap, double &rc) {
        const int64_t k = P.at(i, c);
        const double theta_i = P.Y_theta_i[k];
        const double nu = __ldg(P.nu + k), rcds = __ldg(P.rho_c_ds + k);
        const double T = eh_temperature(U1.get(i), U2.get(i), theta_i, nu, rcds, E);
        eK = volumetric_internal_energy_liq(T, E) * __ldg(P.K_lag + k);
        kap = __ldg(P.kappa_lag + k);
        // Jacobian uses the LAGGED theta_l for rho_c_s (energy_hydrology.jl:561-566)
        rc = 1 / volumetric_heat_capacity(__ldg(P.theta_l_lag + k), theta_i, rcds, E);
    };
    double eKa, kapa, rc0, rc_m = 0.0;
    level2(0, eKa, kapa, rc0);
    double aE_lo = 0.0, aC_lo = 0.0, y_m = 0.0;
    cprev = 0.0; dprev = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double eKb = 0, kapb = 0, rc1 = 0, aE_hi = 0.0, aC_hi = 0.0, y_p = 0.0;
        if (i < N - 1) {
            level2(i + 1, eKb, kapb, rc1);
            const double idzf = G.idzf(i + 1);
            aE_hi = ((eKa + eKb) / 2.0) * idzf;
            aC_hi = ((kapa + kapb) / 2.0) * idzf;
            y_p = D.get(i + 1);
        }
        const double idzc = G.idzc(i);
        const double y0 = D.get(i), x1 = dp.get(i);
        // (W21 x1)_i with W21 = -dtg*(D . Diag(interp(-eK)) . G . Diag(dpsi)) - I
        const double s = dtg * ((aE_lo * (y_m - y0) + aE_hi * (y_p - y0)) * idzc) - x1;
        const double b2 = F2.get(i) - s;
        double lo, di, up;
        tridiag_row(dtg, aC_lo, aC_hi, rc_m, rc0, rc1, idzc, 0.0, lo, di, up);
        const double den = 1.0 / (di - lo * cprev);
        cprev = up * den;
        dprev = (b2 - lo * dprev) * den;
        cp.set(i, cprev);
        F2.set(i, dprev);
        aE_lo = aE_hi; aC_lo = aC_hi; y_m = y0; rc_m = rc0; rc0 = rc1;
        eKa = eKb; kapa = kapb;
    }
    // ---- back substitution 2 and update of both fields
    x = dprev;
    dx2 += x * x;
    U2.set(N - 1, U2.get(N - 1) - x);
    U1.set(N - 1, U1.get(N - 1) - dp.get(N - 1));
#pragma unroll
    for (int i = N - 2; i >= 0; --i) {
        x = F2.get(i) - cp.get(i) * x;
        dx2 += x * x;
        U2.set(i, U2.get(i) - x);
        U1.set(i, U1.get(i) - dp.get(i));
    }
    // flux integrals (W = -I)
    double Tiw = -(top_w - bot_w), Tie = -(top_h - bot_h);
    if (P.topmodel) {
        Tiw -= R_ss;
        Tie -= R_ess;
    }
    const double dxw = -(temp_intw + dtg * Tiw - Uintw);
    const double dxe = -(temp_inte + dtg * Tie - Uinte);
    Uintw -= dxw;
    Uinte -= dxe;
    return dx2 + dxw * dxw + dxe * dxe;
}

template <int CLOSURE, int MATH, int NS>
__global__ void __launch_bounds__(128) k_eh_step_reg(const DevView P, const GridConst<NS> gc, double dtg,
                                                     int max_iters)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < P.ncol;
    double dx2 = 0.0, bad = 0.0;
    if (live) {
        const GridS<NS> G{gc};
        RegCol<NS> U1, U2, cp, dp, D, F2;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            U1.set(i, P.Y_theta_l[P.at(i, c)]);
            U2.set(i, P.Y_rho_e[P.at(i, c)]);
        }
        const double tw = P.Y_intF_w[c], te = P.Y_intF_e[c];
        double Uw = tw, Ue = te;
#pragma unroll 1
        for (int it = 0; it < max_iters; ++it)
            dx2 = eh_newton_iteration<CLOSURE, MATH, NS>(P, G, c, dtg, U1, U2, cp, dp, D, F2, it == 0, Uw, Ue, tw, te);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double a = U1.get(i), b = U2.get(i);
            P.out_theta_l[P.at(i, c)] = a;
            P.out_rho_e[P.at(i, c)] = b;
            if (!isfinite(a)) bad += 1.0;
            if (!isfinite(b)) bad += 1.0;
        }
        P.out_intF_w[c] = Uw;
        P.out_intF_e[c] = Ue;
    }
    accumulate_stats(P, dx2, bad);
}

template <int CLOSURE, int MATH>
__global__ void __launch_bounds__(128) k_eh_step_generic(const DevView P, double dtg, int iter_begin, int iter_end)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < P.ncol && !(P.converged && *P.converged);
    double dx2 = 0.0;
    if (live) {
        const GridR G{P.z_c, P.inv_dz_c, P.inv_dz_f};
        const int N = P.N;
        const int64_t o = P.at(0, c);
        MemCol U1{P.work[0] + o, P.sl}, U2{P.work[1] + o, P.sl}, cp{P.work[2] + o, P.sl}, dp{P.work[3] + o, P.sl};
        MemCol D{P.work[4] + o, P.sl}, F2{P.work[5] + o, P.sl};
        double *carry = P.carry + c;
        const double tw = P.Y_intF_w[c], te = P.Y_intF_e[c];
        double Uw, Ue;
        if (iter_begin == 0) {
            for (int i = 0; i < N; ++i) {
                U1.set(i, P.Y_theta_l[P.at(i, c)]);
                U2.set(i, P.Y_rho_e[P.at(i, c)]);
            }
            Uw = tw;
            Ue = te;
        } else {
            Uw = carry[0];
            Ue = carry[P.ld];
        }
        for (int it = iter_begin; it < iter_end; ++it)
            dx2 = eh_newton_iteration<CLOSURE, MATH, 0>(P, G, c, dtg, U1, U2, cp, dp, D, F2, it == 0, Uw, Ue, tw, te);
        carry[0] = Uw;
        carry[P.ld] = Ue;
    }
    accumulate_stats(P, dx2, 0.0);
}

// Tolerance path, after each iteration (and after the all-reduce of stats[0]):
// converged <- ||dx||_2 <= tol ; the norm of this iteration is kept, the accumulator reset.
__global__ void k_convergence_test(const DevView P, double tol, double *norm_out, int32_t *iters_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (!*P.converged) {
            const double nrm = sqrt(P.stats[0]);
            *norm_out = nrm;
            *iters_out += 1;
            if (nrm <= tol) *P.converged = 1;
        }
        P.stats[0] = 0.0;
    }
}

}  // namespace clb
