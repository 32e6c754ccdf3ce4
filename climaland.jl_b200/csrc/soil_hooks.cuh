// soil_hooks.cuh -- the fine-grained hooks as one kernel each (thread per column,
// runtime N, column-fastest fields).  These are the parity-checkable entry points;
// the performance path is the fused stage in soil_fused.cuh.
#pragma once
#include "soil_device.cuh"

namespace clb {

// One column of update_boundary_fluxes! for the water equation.
//   rre.jl:111-149 -> boundary_flux!: boundary_conditions.jl:227-267 (MoistureStateBC top),
//   :282-325 (MoistureStateBC bottom), :340-353 (FreeDrainage), set_dfluxBCdY! :375-411,
//   diffusive_flux shared_utilities/boundary_conditions.jl:63-65.
// K_top/K_bot, psi_top/psi_bot are the end-level values of the column.
template <int CLOSURE, int MATH>
__device__ __forceinline__ void column_boundary_fluxes(const DevView &P, int64_t c, double K_bot, double psi_bot,
                                                       double K_top, double psi_top, double theta_top)
{
    const int N = P.N;
    if (P.top_bc == 1 /*MoistureStateBC*/) {
        const int64_t k = P.at(N - 1, c);
        const HydroCell cell = load_cell(P, k);
        const double dz = P.dz_top;
        // boundary_flux! evaluates psi_bc with nu (not nu - theta_i) for either model
        const double psi_bc = pressure_head<CLOSURE, MATH>(cell, P.theta_bc_top[c], cell.nu);
        P.top_bc_w[c] = -K_top * ((psi_bc + dz) - psi_top) / dz;
        if (P.model == 0 && P.dfluxBCdY)
            P.dfluxBCdY[c] = K_top * dpsidtheta<CLOSURE, MATH>(cell, theta_top, cell.nu) / dz;
    }
    if (P.bottom_bc == 1 /*FreeDrainage*/) {
        P.bot_bc_w[c] = -1 * K_bot;
    } else if (P.bottom_bc == 2 /*MoistureStateBC*/) {
        const HydroCell cell = load_cell(P, P.at(0, c));
        const double dz = P.dz_bot;
        const double psi_bc = pressure_head<CLOSURE, MATH>(cell, P.theta_bc_bot[c], cell.nu);
        P.bot_bc_w[c] = -K_bot * ((psi_bot + dz) - psi_bc) / dz;
    }
}

// update_implicit_cache!: models.jl:238-246 (aux, then boundary fluxes).
//   Richards  rre.jl:368-380 (K, psi, total_water) + :460-468 (BCs only if dfluxBCdY is cached)
//   EH        energy_hydrology.jl:427-455 (T, psi; K, kappa, theta_l and BCs stay lagged)
template <int CLOSURE, int MATH>
__global__ void __launch_bounds__(128) k_update_implicit_cache(const DevView P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const int N = P.N;
    if (P.model == 0) {
        double tw = 0.0, K_bot = 0.0, psi_bot = 0.0, K = 0.0, psi = 0.0, theta = 0.0;
        for (int i = 0; i < N; ++i) {
            const int64_t k = P.at(i, c);
            const HydroCell cell = load_cell(P, k);
            theta = P.Y_theta_l[k];
            double d;
            CellEval<CLOSURE, MATH>(cell, cell.nu).template eval<true, true, false>(theta, K, psi, d);
            P.p_K[k] = K;
            P.p_psi[k] = psi;
            tw += theta * P.dz_c[i];
            if (i == 0) {
                K_bot = K;
                psi_bot = psi;
            }
        }
        if (P.total_water) P.total_water[c] = tw;
        if (P.top_bc == 1) column_boundary_fluxes<CLOSURE, MATH>(P, c, K_bot, psi_bot, K, psi, theta);
    } else {
        for (int i = 0; i < N; ++i) {
            const int64_t k = P.at(i, c);
            const HydroCell cell = load_cell(P, k);
            const double theta_i = P.Y_theta_i[k];
            const double theta = P.Y_theta_l[k];
            P.p_T[k] = eh_temperature(theta, P.Y_rho_e[k], theta_i, cell.nu, __ldg(P.rho_c_ds + k), P.earth);
            P.p_psi[k] = pressure_head<CLOSURE, MATH>(cell, theta, cell.nu - theta_i);
        }
    }
}

// Explicit-stage flavour: always evaluate the state-type boundary fluxes from the cached K, psi.
template <int CLOSURE, int MATH>
__global__ void __launch_bounds__(128) k_update_boundary_fluxes(const DevView P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const double *K = (P.model == 1) ? P.K_lag : P.p_K;
    const int64_t kb = P.at(0, c), kt = P.at(P.N - 1, c);
    column_boundary_fluxes<CLOSURE, MATH>(P, c, K[kb], P.p_psi[kb], K[kt], P.p_psi[kt], P.Y_theta_l[kt]);
}

// compute_imp_tendency!: rre.jl:161-203, energy_hydrology.jl:363-425.
// InterpolateC2F = mean of the two centres, GradientC2F = centre difference over the
// centre spacing, DivergenceF2C(SetValue) = face-flux difference over the cell thickness
// (test/standalone/Soil/soiltest.jl:357-406).  Implicit source: Runoff/Runoff.jl:321-359.
// NT > 0 (N <= NT): the column's K, psi (T, kappa, is_saturated) in registers, every load issued before the first use --
// in the rolling form (NT = 0) a level's loads wait behind the stores of the level before, which they may alias: one
// HBM round trip per level.  Same expressions in the same order either way.
template <int NT>
__device__ __forceinline__ void imp_tendency_column(const DevView &P, const int64_t c)
{
    const int N = P.N;
    const bool eh = (P.model == 1);
    const double *Kf = eh ? P.K_lag : P.p_K;
    constexpr int NA = NT > 0 ? NT : 1;
    double Kr[NA], pr[NA], Tr[NA], kr[NA], sr[NA];
    if (NT > 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int64_t k = P.at((i < N) ? i : 0, c);
            Kr[i] = __ldg(Kf + k);
            pr[i] = __ldg(P.p_psi + k);
            Tr[i] = eh ? __ldg(P.p_T + k) : 0.0;
            kr[i] = eh ? __ldg(P.kappa_lag + k) : 0.0;
            sr[i] = P.topmodel ? __ldg(P.is_sat + k) : 0.0;
        }
    }
    auto ld_K = [&](int i, int64_t k) { return NT > 0 ? Kr[NT > 0 ? i : 0] : Kf[k]; };
    auto ld_psi = [&](int i, int64_t k) { return NT > 0 ? pr[NT > 0 ? i : 0] : P.p_psi[k]; };
    auto ld_T = [&](int i, int64_t k) { return NT > 0 ? Tr[NT > 0 ? i : 0] : P.p_T[k]; };
    auto ld_kap = [&](int i, int64_t k) { return NT > 0 ? kr[NT > 0 ? i : 0] : P.kappa_lag[k]; };
    auto ld_sat = [&](int i, int64_t k) { return NT > 0 ? sr[NT > 0 ? i : 0] : P.is_sat[k]; };
    const double top_w = P.top_bc_w[c], bot_w = P.bot_bc_w[c];
    double dintw = -(top_w - bot_w), dinte = 0.0;
    double top_h = 0.0, bot_h = 0.0;
    if (eh) {
        top_h = P.top_bc_h[c];
        bot_h = P.bot_bc_h[c];
        dinte = -(top_h - bot_h);
    }
    double src_w = 0.0, src_e = 0.0;
    if (P.topmodel) {
        const double hg = fmax(P.h_grad[c], kEps);
        src_w = P.R_ss[c] / hg;
        dintw -= P.R_ss[c];
        if (eh) {
            src_e = P.R_ess[c] / hg;
            dinte -= P.R_ess[c];
        }
    }
    P.dY_intF_w[c] = dintw;
    if (eh) P.dY_intF_e[c] = dinte;

    const int64_t k00 = P.at(0, c);
    double K0 = ld_K(0, k00), h0 = ld_psi(0, k00) + P.z_c[0];
    double T0 = 0.0, eK0 = 0.0, kap0 = 0.0;
    if (eh) {
        T0 = ld_T(0, k00);
        eK0 = volumetric_internal_energy_liq(T0, P.earth) * K0;
        kap0 = ld_kap(0, k00);
    }
    double qw_lo = bot_w, qe_lo = bot_h;
#pragma unroll
    for (int i = 0; i < (NT > 0 ? NT : N); ++i) {
        if (NT > 0 && i >= N) break;
        const int64_t k = P.at(i, c);
        double qw_hi, qe_hi = 0.0;
        double K1 = 0, h1 = 0, T1 = 0, eK1 = 0, kap1 = 0;
        if (i < N - 1) {
            const int64_t k1 = k + P.sl;
            const int i1 = (NT > 0 && i + 1 >= NT) ? i : i + 1;  // (never taken: i < N - 1 <= NT - 1; keeps the index in bounds)
            const double idzf = __ldg(P.inv_dz_f + i + 1);
            K1 = ld_K(i1, k1);
            h1 = ld_psi(i1, k1) + __ldg(P.z_c + i + 1);
            const double grad_h = (h1 - h0) * idzf;
            qw_hi = -((K0 + K1) / 2.0) * grad_h;
            if (eh) {
                T1 = ld_T(i1, k1);
                eK1 = volumetric_internal_energy_liq(T1, P.earth) * K1;
                kap1 = ld_kap(i1, k1);
                const double grad_T = (T1 - T0) * idzf;
                qe_hi = -((kap0 + kap1) / 2.0) * grad_T - ((eK0 + eK1) / 2.0) * grad_h;
            }
        } else {
            qw_hi = top_w;
            qe_hi = top_h;
        }
        const double idzc = __ldg(P.inv_dz_c + i);
        double tw = -((qw_hi - qw_lo) * idzc);
        double sat = 0.0;
        if (P.topmodel) {
            sat = ld_sat(i, k);
            tw -= src_w * sat;
        }
        P.dY_theta_l[k] = tw;
        if (eh) {
            double te = -((qe_hi - qe_lo) * idzc);
            if (P.topmodel) te -= src_e * sat;
            P.dY_rho_e[k] = te;
            P.dY_theta_i[k] = 0.0;
        }
        qw_lo = qw_hi;
        qe_lo = qe_hi;
        K0 = K1; h0 = h1; T0 = T1; eK0 = eK1; kap0 = kap1;
    }
}

__global__ void __launch_bounds__(128) k_imp_tendency(const DevView P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    if (P.N <= 16) imp_tendency_column<16>(P, c);  // warp-uniform
    else imp_tendency_column<0>(P, c);
}

// One tridiagonal row of  W = -dtgamma * (D . Diag(interp(-A)) . G . Diag(coef)) - I
// (rre.jl:423-454, energy_hydrology.jl:503-573).  a_lo / a_hi are the face values
// interp(A)/dz_f below / above the cell (0 at the boundary faces: SetGradient(0)).
__device__ __forceinline__ void tridiag_row(double dtg, double a_lo, double a_hi, double coef_m, double coef_0,
                                            double coef_p, double idzc, double top_dflux, double &lo, double &di,
                                            double &up)
{
    lo = dtg * (a_lo * coef_m) * idzc;
    up = dtg * (a_hi * coef_p) * idzc;
    di = -dtg * (((a_hi + a_lo) * coef_0 + top_dflux) * idzc) - 1.0;
}

// compute_jacobian!: rre.jl:391-458, energy_hydrology.jl:466-576.
template <int CLOSURE, int MATH>
__global__ void __launch_bounds__(128) k_jacobian(const DevView P, double dtg)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const int N = P.N;
    const bool eh = (P.model == 1);
    const double *Kf = eh ? P.K_lag : P.p_K;
    // haskey(p.soil, :dfluxBCdY): only a Richards cache with a MoistureStateBC top has it
    const double top_dflux = (!eh && P.top_bc == 1) ? P.dfluxBCdY[c] : 0.0;

    // rolling window over levels i-1, i, i+1
    double dps_m = 0, dps_0 = 0, dps_p = 0;   // dpsi/dtheta
    double K_m = 0, K_0 = 0, K_p = 0;
    double eK_m = 0, eK_0 = 0, eK_p = 0;      // e_liq(T) * K
    double kp_m = 0, kp_0 = 0, kp_p = 0;      // kappa
    double rc_m = 0, rc_0 = 0, rc_p = 0;      // 1 / rho_c_s(lagged theta_l, theta_i)

    auto level = [&](int i, double &dps, double &K, double &eK, double &kp, double &rc) {
        const int64_t k = P.at(i, c);
        const HydroCell cell = load_cell(P, k);
        const double theta_i = eh ? P.Y_theta_i[k] : 0.0;
        dps = dpsidtheta<CLOSURE, MATH>(cell, P.Y_theta_l[k], eh ? cell.nu - theta_i : cell.nu);
        K = Kf[k];
        if (eh) {
            eK = volumetric_internal_energy_liq(P.p_T[k], P.earth) * K;
            kp = P.kappa_lag[k];
            rc = 1 / volumetric_heat_capacity(P.theta_l_lag[k], theta_i, __ldg(P.rho_c_ds + k), P.earth);
        }
    };
    level(0, dps_0, K_0, eK_0, kp_0, rc_0);
    double aK_lo = 0, aE_lo = 0, aC_lo = 0;
    for (int i = 0; i < N; ++i) {
        const int64_t k = P.at(i, c);
        double aK_hi = 0, aE_hi = 0, aC_hi = 0;
        if (i < N - 1) {
            level(i + 1, dps_p, K_p, eK_p, kp_p, rc_p);
            const double idzf = P.inv_dz_f[i + 1];
            aK_hi = ((K_0 + K_p) / 2.0) * idzf;
            aE_hi = ((eK_0 + eK_p) / 2.0) * idzf;
            aC_hi = ((kp_0 + kp_p) / 2.0) * idzf;
        } else {
            dps_p = 0; rc_p = 0;
        }
        const double idzc = P.inv_dz_c[i];
        double lo, di, up;
        tridiag_row(dtg, aK_lo, aK_hi, dps_m, dps_0, dps_p, idzc, (i == N - 1) ? top_dflux : 0.0, lo, di, up);
        P.w11_lo[k] = lo; P.w11_di[k] = di; P.w11_up[k] = up;
        if (eh) {
            // (rho_e_int, theta_l): A = e_liq*K, same coef, and "- I" (sic, energy_hydrology.jl:554-556)
            tridiag_row(dtg, aE_lo, aE_hi, dps_m, dps_0, dps_p, idzc, 0.0, lo, di, up);
            P.w21_lo[k] = lo; P.w21_di[k] = di; P.w21_up[k] = up;
            // (rho_e_int, rho_e_int): A = kappa, coef = 1/rho_c_s
            tridiag_row(dtg, aC_lo, aC_hi, rc_m, rc_0, rc_p, idzc, 0.0, lo, di, up);
            P.w22_lo[k] = lo; P.w22_di[k] = di; P.w22_up[k] = up;
        }
        aK_lo = aK_hi; aE_lo = aE_hi; aC_lo = aC_hi;
        dps_m = dps_0; dps_0 = dps_p;
        K_m = K_0; K_0 = K_p;
        eK_m = eK_0; eK_0 = eK_p;
        kp_m = kp_0; kp_0 = kp_p;
        rc_m = rc_0; rc_0 = rc_p;
    }
    (void)K_m; (void)eK_m; (void)kp_m;
}

// Thomas sweep in the normalised (c', d') form; x doubles as d'.
__device__ __forceinline__ void thomas_column(int N, int64_t ld /* level stride */, const double *lo, const double *di, const double *up,
                                              const double *b, double *x, double *cp)
{
    double den = 1.0 / di[0];
    double cprev = up[0] * den, xprev = b[0] * den;
    cp[0] = cprev;
    x[0] = xprev;
    for (int i = 1; i < N; ++i) {
        const int64_t k = (int64_t)i * ld;
        const double l = lo[k];
        den = 1.0 / (di[k] - l * cprev);
        cprev = up[k] * den;
        xprev = (b[k] - l * xprev) * den;
        cp[k] = cprev;
        x[k] = xprev;
    }
    for (int i = N - 2; i >= 0; --i) {
        const int64_t k = (int64_t)i * ld;
        xprev = x[k] - cp[k] * xprev;
        x[k] = xprev;
    }
}

// The same sweep for N <= NT with the column in registers: every load of the column is issued before the first use (in
// the loop above a level's loads wait for the level before: they may alias its stores, and the recurrence is a chain of
// IEEE divisions), c' and d' never touch memory.  Same expressions in the same order: the same bits.  r: b in, x out.
template <int NT>
__device__ __forceinline__ void thomas_regs(int N, int64_t ld, const double *__restrict__ lo, const double *__restrict__ di,
                                            const double *__restrict__ up, double (&r)[NT])
{
    double l[NT], d[NT], u[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int64_t k = (int64_t)((i < N) ? i : 0) * ld;
        l[i] = __ldg(lo + k);
        d[i] = __ldg(di + k);
        u[i] = __ldg(up + k);
    }
    double den = 1.0 / d[0];
    double cprev = u[0] * den, xprev = r[0] * den;
    u[0] = cprev;
    r[0] = xprev;
#pragma unroll
    for (int i = 1; i < NT; ++i) {
        if (i < N) {
            den = 1.0 / (d[i] - l[i] * cprev);
            cprev = u[i] * den;
            xprev = (r[i] - l[i] * xprev) * den;
            u[i] = cprev;
            r[i] = xprev;
        }
    }
#pragma unroll
    for (int i = NT - 2; i >= 0; --i) {
        if (i < N - 1) {
            xprev = r[i] - u[i] * xprev;
            r[i] = xprev;
        }
    }
}

// the soil blocks of ldiv! for one column with N <= NT levels, in registers (see thomas_regs); values as ldiv_soil_column
template <int NT>
__device__ __forceinline__ void ldiv_soil_regs(const DevView &P, int64_t c)
{
    const int N = P.N;
    const int64_t ld = P.sl, o = P.at(0, c);
    auto lev = [&](int i) { return (int64_t)((i < N) ? i : 0) * ld + o; };
    double x1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) x1[i] = __ldg(P.b_theta_l + lev(i));
    thomas_regs<NT>(N, ld, P.w11_lo + o, P.w11_di + o, P.w11_up + o, x1);
#pragma unroll
    for (int i = 0; i < NT; ++i)
        if (i < N) P.x_theta_l[lev(i)] = x1[i];
    P.x_intF_w[c] = -P.b_intF_w[c];
    if (P.model == 1) {
        double b2[NT], nb[NT];
        {
            double wl[NT], wd[NT], wu[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                wl[i] = __ldg(P.w21_lo + lev(i));
                wd[i] = __ldg(P.w21_di + lev(i));
                wu[i] = __ldg(P.w21_up + lev(i));
                b2[i] = __ldg(P.b_rho_e + lev(i));
                nb[i] = __ldg(P.b_theta_i + lev(i));
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                double s = wd[i] * x1[i];
                if (i > 0) s = wl[i] * x1[i - 1] + s;
                if (i < N - 1 && i < NT - 1) s = s + wu[i] * x1[(i < NT - 1) ? i + 1 : i];
                b2[i] = b2[i] - s;
            }
        }
        thomas_regs<NT>(N, ld, P.w22_lo + o, P.w22_di + o, P.w22_up + o, b2);
#pragma unroll
        for (int i = 0; i < NT; ++i)
            if (i < N) {
                P.x_rho_e[lev(i)] = b2[i];
                P.x_theta_i[lev(i)] = -nb[i];
            }
        P.x_intF_e[c] = -P.b_intF_e[c];
    }
}

// ldiv!: implicit_timestepping.jl:160-171.  Richards: BlockDiagonalSolve.  EH:
// BlockLowerTriangularSolve(theta_l): W11 x1 = b1; b2' = b2 - W21 x1; W22 x2 = b2'.
// x = -b for the -I blocks (theta_i and the two flux integrals).
__global__ void __launch_bounds__(128) k_ldiv(const DevView P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const int N = P.N;
    if (N <= 16) {  // warp-uniform
        ldiv_soil_regs<16>(P, c);
        return;
    }
    const int64_t ld = P.sl, o = P.at(0, c);
    double *cp = P.work[0] + o;
    thomas_column(N, ld, P.w11_lo + o, P.w11_di + o, P.w11_up + o, P.b_theta_l + o, P.x_theta_l + o, cp);
    P.x_intF_w[c] = -P.b_intF_w[c];
    if (P.model == 1) {
        double *b2 = P.work[1] + o;
        const double *x1 = P.x_theta_l + o;
        for (int i = 0; i < N; ++i) {
            const int64_t k = (int64_t)i * ld;
            double s = P.w21_di[k + o] * x1[k];
            if (i > 0) s = P.w21_lo[k + o] * x1[k - ld] + s;
            if (i < N - 1) s = s + P.w21_up[k + o] * x1[k + ld];
            b2[k] = P.b_rho_e[k + o] - s;
        }
        thomas_column(N, ld, P.w22_lo + o, P.w22_di + o, P.w22_up + o, b2, P.x_rho_e + o, cp);
        for (int i = 0; i < N; ++i) {
            const int64_t k = P.at(i, c);
            P.x_theta_i[k] = -P.b_theta_i[k];
        }
        P.x_intF_e[c] = -P.b_intF_e[c];
    }
}

// ldiv! of an integrated model's block-diagonal FieldMatrix in one launch (implicit_timestepping.jl:63-172): blockIdx.y =
// 0 the soil blocks (k_ldiv), 1 / 2 the (CO2, CO2) / (O2, O2) tridiagonals, 3 the DiagonalMatrixRow block of one
// surface variable; `blocks` = CLB_LDIV_* mask of what is present.
struct LdivAllView {
    const double *co2_lo[2], *co2_di[2], *co2_up[2], *co2_b[2];
    double *co2_x[2];
    const double *sfc_w, *sfc_b;
    double *sfc_x;
    unsigned blocks;
};
__global__ void __launch_bounds__(128) k_ldiv_all(const DevView P, const LdivAllView A)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const int N = P.N, part = blockIdx.y;
    const int64_t ld = P.sl, o = P.at(0, c);
    if (part == 0) {
        if (!(A.blocks & 1u)) return;
        if (N <= 16) {
            ldiv_soil_regs<16>(P, c);
            return;
        }
        double *cp = P.work[0] + o;
        thomas_column(N, ld, P.w11_lo + o, P.w11_di + o, P.w11_up + o, P.b_theta_l + o, P.x_theta_l + o, cp);
        P.x_intF_w[c] = -P.b_intF_w[c];
        if (P.model == 1) {
            double *b2 = P.work[1] + o;
            const double *x1 = P.x_theta_l + o;
            for (int i = 0; i < N; ++i) {
                const int64_t k = (int64_t)i * ld;
                double s = P.w21_di[k + o] * x1[k];
                if (i > 0) s = P.w21_lo[k + o] * x1[k - ld] + s;
                if (i < N - 1) s = s + P.w21_up[k + o] * x1[k + ld];
                b2[k] = P.b_rho_e[k + o] - s;
            }
            thomas_column(N, ld, P.w22_lo + o, P.w22_di + o, P.w22_up + o, b2, P.x_rho_e + o, cp);
            for (int i = 0; i < N; ++i) {
                const int64_t k = P.at(i, c);
                P.x_theta_i[k] = -P.b_theta_i[k];
            }
            P.x_intF_e[c] = -P.b_intF_e[c];
        }
    } else if (part <= 2) {
        if (!(A.blocks & 2u)) return;
        const int sp = part - 1;
        if (N <= 16) {
            double r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __ldg(A.co2_b[sp] + o + (int64_t)((i < N) ? i : 0) * ld);
            thomas_regs<16>(N, ld, A.co2_lo[sp] + o, A.co2_di[sp] + o, A.co2_up[sp] + o, r);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < N) A.co2_x[sp][o + (int64_t)i * ld] = r[i];
            return;
        }
        thomas_column(N, ld, A.co2_lo[sp] + o, A.co2_di[sp] + o, A.co2_up[sp] + o, A.co2_b[sp] + o, A.co2_x[sp] + o,
                      P.work[2 + sp] + o);
    } else {
        if (!(A.blocks & 4u)) return;
        A.sfc_x[c] = A.sfc_b[c] / A.sfc_w[c];
    }
}

// out[c] = sum_i field[i,c]*dz_c[i]   (column_integral_definite!, rre.jl:502-511)
__global__ void __launch_bounds__(128) k_column_integral(const DevView P, const double *field, double *out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    double s = 0.0;
    for (int i = 0; i < P.N; ++i) s += field[P.at(i, c)] * P.dz_c[i];
    out[c] = s;
}

}  // namespace clb
