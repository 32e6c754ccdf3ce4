// soil_co2.cuh -- SoilCO2Model's implicit CO2 / O2 diffusion: two more independent per-column
// tridiagonals with the same D . Diag(interp) . G . Diag(coef) structure as the soil water equation, with lagged
// coefficients (p.soilco2.{D, theta_eff} and {D_o2, theta_eff_o2}).
//
// Reference (paths relative to the ClimaLand.jl tree, src/standalone/Soil/Biogeochemistry/Biogeochemistry.jl):
//   make_update_implicit_boundary_fluxes   :320-357  (AtmosCO2StateBC :932-957, AtmosO2StateBC :1078-1111,
//                                                     diffusive_flux shared_utilities/boundary_conditions.jl:63-65)
//   make_compute_imp_tendency              :371-413
//   make_compute_jacobian                  :1119-1195
//   Newton / ARS111 stage                  src/simulations/Simulations.jl:127-135
//
// One thread per (column, species): blockIdx.y = 0 CO2, 1 O2.  No transcendental functions: the kernels stream
// 4 cell fields in and 1 out per species (HBM-bound).  N <= kCo2MaxLevels; Thomas vectors in local memory.
#pragma once
#include "soil_device.cuh"

namespace clb {

constexpr int kCo2MaxLevels = 64;

struct Co2Species {
    double *C;                        // state (in / out of the stage)
    const double *D, *theta_eff;      // lagged cell fields
    double *top_bc;                   // per column; written when the top BC is the atmosphere's state
    const double *bot_bc, *c_atm;     // per column; c_atm == nullptr: top_bc is a flux value
    double *dflux;                    // dfluxBCdY, per column (state BC only)
    double *dC;                       // implicit tendency
    double *lo, *di, *up;             // Jacobian rows
};
struct Co2View {
    Co2Species s[2];
};

// boundary_flux! of the state BC for one column; returns dfluxBCdY (0 for a flux BC)
__device__ __forceinline__ double co2_top_flux(const DevView &P, const Co2Species &S, int64_t c, double C_top, double &top)
{
    if (!S.c_atm) {
        top = S.top_bc[c];
        return 0.0;
    }
    const int64_t q = P.at(P.N - 1, c);
    const double D = S.D[q], th = S.theta_eff[q];
    top = -D * (S.c_atm[c] - fmax(C_top / th, 0.0)) / P.dz_top;
    return D / th / P.dz_top;
}

// MODE 0: boundary fluxes only, 1: implicit tendency, 2: Jacobian rows, 3: the fused stage
template <int MODE, int NS>
__global__ void __launch_bounds__(128) k_co2(const DevView P, const Co2View V, double dtg, int max_iters)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.ncol) return;
    const Co2Species &S = V.s[blockIdx.y];
    const int N = (NS > 0) ? NS : P.N;
    if (MODE == 0) {
        double top;
        const double d = co2_top_flux(P, S, c, S.C[P.at(N - 1, c)], top);
        if (S.c_atm) {
            S.top_bc[c] = top;
            S.dflux[c] = d;
        }
        return;
    }
    if (MODE == 1) {
        const double top = S.top_bc[c];
        double q_lo = S.bot_bc[c];
        double u0 = fmax(S.C[P.at(0, c)], 0.0) / S.theta_eff[P.at(0, c)], D0 = S.D[P.at(0, c)];
        for (int i = 0; i < N; ++i) {
            double q_hi = top, u1 = 0.0, D1 = 0.0;
            if (i < N - 1) {
                const int64_t q1 = P.at(i + 1, c);
                u1 = fmax(S.C[q1], 0.0) / S.theta_eff[q1];
                D1 = S.D[q1];
                q_hi = -((D0 + D1) / 2.0) * ((u1 - u0) * P.inv_dz_f[i + 1]);
            }
            S.dC[P.at(i, c)] = -((q_hi - q_lo) * P.inv_dz_c[i]);
            q_lo = q_hi;
            u0 = u1;
            D0 = D1;
        }
        return;
    }
    if (MODE == 2) {
        const double dflux = S.c_atm ? S.dflux[c] : 0.0;
        double Dm = 0.0, D0 = S.D[P.at(0, c)], rm = 0.0, r0 = 1.0 / S.theta_eff[P.at(0, c)];
        for (int i = 0; i < N; ++i) {
            double Dp = 0.0, rp = 0.0;
            if (i < N - 1) {
                Dp = S.D[P.at(i + 1, c)];
                rp = 1.0 / S.theta_eff[P.at(i + 1, c)];
            }
            const double a_lo = (i > 0) ? ((Dm + D0) / 2.0) * P.inv_dz_f[i] : 0.0;
            const double a_hi = (i < N - 1) ? ((D0 + Dp) / 2.0) * P.inv_dz_f[i + 1] : 0.0;
            const double idz = P.inv_dz_c[i];
            const int64_t q = P.at(i, c);
            S.lo[q] = dtg * (a_lo * rm) * idz;
            S.up[q] = dtg * (a_hi * rp) * idz;
            S.di[q] = -dtg * (((a_hi + a_lo) * r0 + ((i == N - 1) ? dflux : 0.0)) * idz) - 1.0;
            Dm = D0; D0 = Dp; rm = r0; r0 = rp;
        }
        return;
    }
    // ---- fused stage: max_iters x (boundary flux, Jacobian, tendency, residual, Thomas, update) ----
    // D and theta_eff are lagged, so W = dtgamma dT/dC - I (including dfluxBCdY = D_N / theta_N / dz_top) is the
    // same matrix in every Newton iteration: it is built and factored once (Thomas' c' and 1/pivot), and an
    // iteration is the tendency stencil, one forward and one backward sweep.  NS > 0: the column lives in registers.
    if constexpr (NS == 0) {
        // any N <= kCo2MaxLevels: iterate and Thomas vectors in local memory (4 arrays), the lagged fields re-read
        // from L1 / L2 in every iteration with rolling neighbours
        double U[kCo2MaxLevels], den[kCo2MaxLevels], cp[kCo2MaxLevels], g[kCo2MaxLevels];
        const double bot = S.bot_bc[c];
        const int64_t qN = P.at(N - 1, c);
        const double D_top = S.D[qN], r_top = 1.0 / S.theta_eff[qN];
        const double dflux = S.c_atm ? D_top * r_top / P.dz_top : 0.0;
        {
            double Dm = 0.0, D0 = S.D[P.at(0, c)], rm = 0.0, r0 = 1.0 / S.theta_eff[P.at(0, c)];
            for (int i = 0; i < N; ++i) {
                U[i] = S.C[P.at(i, c)];
                double Dp = 0.0, rp = 0.0;
                if (i < N - 1) {
                    Dp = S.D[P.at(i + 1, c)];
                    rp = 1.0 / S.theta_eff[P.at(i + 1, c)];
                }
                const double a_lo = (i > 0) ? ((Dm + D0) / 2.0) * P.inv_dz_f[i] : 0.0;
                const double a_hi = (i < N - 1) ? ((D0 + Dp) / 2.0) * P.inv_dz_f[i + 1] : 0.0;
                const double idz = P.inv_dz_c[i];
                const double lo = dtg * (a_lo * rm) * idz, up = dtg * (a_hi * rp) * idz;
                const double di = -dtg * (((a_hi + a_lo) * r0 + ((i == N - 1) ? dflux : 0.0)) * idz) - 1.0;
                den[i] = 1.0 / (di - ((i > 0) ? lo * cp[i - 1] : 0.0));
                cp[i] = up * den[i];
                Dm = D0; D0 = Dp; rm = r0; r0 = rp;
            }
        }
        double top = S.c_atm ? 0.0 : S.top_bc[c];
        for (int it = 0; it < max_iters; ++it) {
            if (S.c_atm) top = -D_top * (S.c_atm[c] - fmax(U[N - 1] * r_top, 0.0)) / P.dz_top;
            double Dm = 0.0, D0 = S.D[P.at(0, c)], rm = 0.0, r0 = 1.0 / S.theta_eff[P.at(0, c)];
            double q_lo = bot, u0 = fmax(U[0], 0.0) * r0;
            for (int i = 0; i < N; ++i) {
                double Dp = 0.0, rp = 0.0, u1 = 0.0;
                if (i < N - 1) {
                    Dp = S.D[P.at(i + 1, c)];
                    rp = 1.0 / S.theta_eff[P.at(i + 1, c)];
                    u1 = fmax(U[i + 1], 0.0) * rp;
                }
                const double a_lo = (i > 0) ? ((Dm + D0) / 2.0) * P.inv_dz_f[i] : 0.0;
                const double a_hi = (i < N - 1) ? ((D0 + Dp) / 2.0) * P.inv_dz_f[i + 1] : 0.0;
                const double idz = P.inv_dz_c[i];
                const double q_hi = (i < N - 1) ? -(a_hi * (u1 - u0)) : top;
                const double f = S.C[P.at(i, c)] + dtg * (-((q_hi - q_lo) * idz)) - U[i];
                const double lo = dtg * (a_lo * rm) * idz;
                g[i] = (f - ((i > 0) ? lo * g[i - 1] : 0.0)) * den[i];
                q_lo = q_hi;
                u0 = u1;
                Dm = D0; D0 = Dp; rm = r0; r0 = rp;
            }
            double x = g[N - 1];
            U[N - 1] -= x;
            for (int i = N - 2; i >= 0; --i) {
                x = g[i] - cp[i] * x;
                U[i] -= x;
            }
        }
        if (S.c_atm) S.top_bc[c] = top;
        for (int i = 0; i < N; ++i) S.C[P.at(i, c)] = U[i];
        return;
    }
    constexpr int NA = (NS > 0) ? NS : kCo2MaxLevels;
    double U[NA], r[NA], a[NA + 1], den[NA], cp[NA], g[NA];
    {
        double Dm = 0.0;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (i < N) {
                const int64_t q = P.at(i, c);
                const double D0 = S.D[q];
                U[i] = S.C[q];
                r[i] = 1.0 / S.theta_eff[q];
                a[i] = (i > 0) ? ((Dm + D0) / 2.0) * P.inv_dz_f[i] : 0.0;  // face below cell i
                Dm = D0;
            }
        }
        a[N] = 0.0;
    }
    const double bot = S.bot_bc[c];
    const double D_top = S.D[P.at(N - 1, c)];
    const double dflux = S.c_atm ? D_top * r[N - 1] / P.dz_top : 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (i < N) {
            const double idz = P.inv_dz_c[i];
            const double lo = (i > 0) ? dtg * (a[i] * r[i - 1]) * idz : 0.0;
            const double up = (i < N - 1) ? dtg * (a[i + 1] * r[i + 1]) * idz : 0.0;
            const double di = -dtg * (((a[i + 1] + a[i]) * r[i] + ((i == N - 1) ? dflux : 0.0)) * idz) - 1.0;
            den[i] = 1.0 / (di - ((i > 0) ? lo * cp[i - 1] : 0.0));
            cp[i] = up * den[i];
        }
    }
    double top = S.c_atm ? 0.0 : S.top_bc[c];
    for (int it = 0; it < max_iters; ++it) {
        if (S.c_atm) top = -D_top * (S.c_atm[c] - fmax(U[N - 1] * r[N - 1], 0.0)) / P.dz_top;
        double q_lo = bot, u0 = fmax(U[0], 0.0) * r[0];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (i < N) {
                const double u1 = (i < N - 1) ? fmax(U[i + 1], 0.0) * r[i + 1] : 0.0;
                const double q_hi = (i < N - 1) ? -(a[i + 1] * (u1 - u0)) : top;
                const double idz = P.inv_dz_c[i];
                const double temp = S.C[P.at(i, c)];  // the stage's `temp` is the untouched input (an L1 / L2 hit)
                const double f = temp + dtg * (-((q_hi - q_lo) * idz)) - U[i];
                const double lo = (i > 0) ? dtg * (a[i] * r[i - 1]) * idz : 0.0;
                g[i] = (f - ((i > 0) ? lo * g[i - 1] : 0.0)) * den[i];
                q_lo = q_hi;
                u0 = u1;
            }
        }
        double x = g[N - 1];
        U[N - 1] -= x;
#pragma unroll
        for (int i = NA - 2; i >= 0; --i) {
            if (i < N - 1) {
                x = g[i] - cp[i] * x;
                U[i] -= x;
            }
        }
    }
    if (S.c_atm) S.top_bc[c] = top;  // the cache keeps the flux of the last evaluation, as the reference's does
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if (i < N) S.C[P.at(i, c)] = U[i];
}

}  // namespace clb
