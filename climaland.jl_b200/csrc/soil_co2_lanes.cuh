// soil_co2_lanes.cuh -- the fused SoilCO2 stage (soil_co2.cuh, MODE 3) on the lane geometry of the soil lane
// kernels (soil_pair.cuh: LaneGeom): a (column, species) system is split over 2 * PARTS lanes of a warp, Q cells per
// lane ordered from the column's boundary towards the seam, 32 / (2 PARTS) columns per warp.
//
// Why: one thread per (column, species) keeps 6 N doubles of one system in registers and runs its sweeps as one
// serial chain, so a ~1 degree domain (61 206 columns) puts 6.5 warps on an SM and the stage is bound by the latency
// of those chains (58 us, 0.16 of the HBM roofline; profiles/r1/soilco2_stage_timing.txt).  Here the same stage has
// 2 PARTS x the threads, the stencil needs two lane crossings per iteration, and the tridiagonal solve is the twisted
// (two-sided) Thomas factorisation of the soil kernels: W = dtgamma dT/dC - I is built from lagged fields only
// (D, theta_eff; dfluxBCdY = D_N / theta_N / dz_top), so it is factored ONCE per stage -- elimination on the leading
// minors, boundary -> seam, the pivots' reciprocals afterwards and in parallel -- and a Newton iteration is the
// tendency stencil, one forward and one backward substitution per half and the 2 x 2 seam system.
//
// Reference (src/standalone/Soil/Biogeochemistry/Biogeochemistry.jl): update_implicit_boundary_fluxes :320-357
// (AtmosCO2StateBC :932-957, AtmosO2StateBC :1078-1111), compute_imp_tendency :371-413, compute_jacobian :1119-1195;
// Newton / ARS111 stage src/simulations/Simulations.jl:127-135.
//
// Inward flux convention as in soil_pair.cuh: F_f = -a_f (u_inner - u_outer) [+ boundary flux] is the flux through
// face f in the direction boundary -> seam, T_q = (F_q - F_{q+1}) / dz_q, i.e. the reference's -(q_hi - q_lo)/dz with
// exact sign flips.  Divisions by theta_eff and by the pivots are the branch-free reciprocals of soil_math.cuh
// (<= 1 ulp from the IEEE quotient); CLB_MATH_LIBM handles keep the thread-per-column kernel and its IEEE divisions.
#pragma once
#include "soil_co2.cuh"
#include "soil_pair.cuh"

namespace clb {

#ifndef CLB_CO2_MINB
#define CLB_CO2_MINB 5  // <= 102 registers, 5 blocks per SM: 26.8 us at ~1 degree against 28.8 (4 blocks), 28.0 (6), 35.2 (8: spills)
#endif
template <int PARTS, int Q>
__global__ void __launch_bounds__(128, (Q <= 4) ? CLB_CO2_MINB : 1) k_co2_lanes(const DevView P, const Co2View V, double dtg, int max_iters)
{
    using Gm = LaneGeom<PARTS, Q>;
    constexpr int CPW = Gm::CPW, NR = Gm::NR;
    const Co2Species &S = V.s[blockIdx.y];
    const int N = P.N;  // NR / 2 < N <= NR (the host picks PARTS and Q)
    const int lane = threadIdx.x & 31;
    const int idx = lane / CPW, half = idx & 1, part = idx >> 1, r0 = part * Q;
    const bool innermost = (part == PARTS - 1), outermost = (part == 0);
    const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t c = tile * CPW + (lane % CPW);
    const bool col_ok = c < P.ncol;
    const int64_t cs = col_ok ? c : P.ncol - 1;  // lanes past the last column work on a copy of it and store nothing
    // the NR - N pad rows sit at the outer end of the top half: its first real slot is slot qT of part pT
    const int Q0T = NR - N, pT = Q0T / Q, qT = Q0T % Q;
    const bool top_lane = half && part == pT;
    auto level_of = [&](int q) { return half ? NR - 1 - (r0 + q) : r0 + q; };
    // 1/dz_f of the face between levels f - 1 and f; zero for the column's boundaries and for faces of pad rows
    auto fidz = [&](int f) { return (f >= 1 && f <= N - 1) ? __ldg(P.inv_dz_f + f) : 0.0; };

    double U[Q], tmp[Q], r[Q], a[Q + 1], dti[Q];
    {
        double D[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int l = level_of(q);
            const bool real = l < N;
            const int64_t k = P.at(real ? l : N - 1, cs);
            const double d_ = S.D[k], th_ = S.theta_eff[k], c_ = S.C[k];
            D[q] = real ? d_ : 0.0;
            U[q] = real ? c_ : 0.0;
            r[q] = fm::rcp(real ? th_ : 1.0);
            dti[q] = real ? dtg * __ldg(P.inv_dz_c + l) : 0.0;
            tmp[q] = U[q];
        }
        double D_out, D_in;
        nb_exchange<Gm>(D[0], D[Q - 1], innermost, D_out, D_in);
#pragma unroll
        for (int q = 0; q <= Q; ++q) {
            // face q of the lane: outer face of cell q (q < Q) / inner face of the last cell (q = Q)
            const int lq = level_of((q < Q) ? q : Q - 1);
            const int f = (q < Q) ? (half ? lq + 1 : lq) : (half ? lq : lq + 1);
            const double Do = (q == 0) ? D_out : D[q - 1], Di = (q < Q) ? D[q] : D_in;
            a[q] = ((Do + Di) / 2.0) * fidz(f);
        }
    }
    // boundary fluxes in the inward convention: bottom_bc enters at face 0 of the bottom half's outer lane, -top_bc at
    // face qT of the top lane
    const double bot = S.bot_bc[cs];
    const bool state_bc = S.c_atm != nullptr;
    const double c_atm = state_bc ? S.c_atm[cs] : 0.0;
    double top = state_bc ? 0.0 : S.top_bc[cs];
    const double inv_dz_top = fm::rcp(P.dz_top);  // (P.inv_dz_top from the host here: 26.9 -> 28.8 us per ~1 degree stage -- the register allocation it leads to)
    double D_top = 0.0, r_top = 1.0;  // the top cell's, in the top lane
    if (state_bc) {
        const int64_t kN = P.at(N - 1, cs);
        D_top = S.D[kN];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (q == qT) r_top = r[q];
    }
    const double dflux = (state_bc && top_lane) ? (D_top * r_top) * inv_dz_top : 0.0;

    // ---- W = dtgamma dT/dC - I, factored once: rows in the lane-local orientation (o: towards the boundary,
    //      i: towards the seam), elimination boundary -> seam on the leading minors (soil_pair.cuh, W22 set-up)
    double od[Q], den[Q], cp[Q], c_last, r_seam;
    {
        double r_out, r_in, o[Q], in_[Q], d[Q];
        nb_exchange<Gm>(r[0], r[Q - 1], innermost, r_out, r_in);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            o[q] = (a[q] * ((q == 0) ? r_out : r[q - 1])) * dti[q];
            in_[q] = (a[q + 1] * ((q < Q - 1) ? r[q + 1] : r_in)) * dti[q];
            d[q] = fma(-((a[q + 1] + a[q]) * r[q] + ((q == qT) ? dflux : 0.0)), dti[q], -1.0);
        }
        double Din = 1.0, iDin = 0.0, Dm[Q], iD[Q];
#pragma unroll
        for (int pass = 0; pass < PARTS; ++pass) {
            double Dp = Din, iDp = iDin;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                iD[q] = in_[q] * Dp;
                Dm[q] = fma(d[q], Dp, -(o[q] * iDp));
                iDp = iD[q];
                Dp = Dm[q];
            }
            if (pass + 1 < PARTS) {
                const double rD_ = from_prev_part<Gm::PARTD>(Dm[Q - 1]), riD_ = from_prev_part<Gm::PARTD>(iD[Q - 1]);
                Din = outermost ? 1.0 : rD_;
                iDin = outermost ? 0.0 : riD_;
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double rD = fm::rcp(Dm[q]);
            den[q] = ((q == 0) ? Din : Dm[q - 1]) * rD;
            od[q] = o[q] * den[q];
            cp[q] = iD[q] * rD;
        }
        c_last = cp[Q - 1];
        r_seam = fm::rcp(fma(-c_last, xchg<Gm::SEAM>(c_last), 1.0));
    }

    const double b0 = (half == 0 && outermost) ? bot : 0.0;
#pragma unroll 1
    for (int it = 0; it < max_iters; ++it) {
        // update_implicit_boundary_fluxes!: diffusive_flux(D_N, c_atm, max(C_N / theta_N, 0), dz_top) at the iterate
        if (state_bc) {
            double U_top = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (q == qT) U_top = U[q];
            top = (-D_top * (c_atm - fmax(U_top * r_top, 0.0))) * inv_dz_top;
        }
        const double bT = top_lane ? -top : 0.0;
        double u[Q], u_out, u_in;
#pragma unroll
        for (int q = 0; q < Q; ++q) u[q] = fmax(U[q], 0.0) * r[q];
        nb_exchange<Gm>(u[0], u[Q - 1], innermost, u_out, u_in);
        // residual f = temp + dtgamma T(U) - U, scaled by the pivots' reciprocals, and the forward substitution
        double g[Q], gin = 0.0, fsc[Q];
        {
            double F_o = ((qT == 0) ? (b0 + bT) : b0) - a[0] * (u[0] - u_out);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double un = (q < Q - 1) ? u[q + 1] : u_in;
                double F_in = -(a[q + 1] * (un - u[q]));
                if (q + 1 == qT) F_in += bT;
                const double f = fma(F_o - F_in, dti[q], tmp[q]) - U[q];
                fsc[q] = f * den[q];
                F_o = F_in;
            }
        }
#pragma unroll
        for (int pass = 0; pass < PARTS; ++pass) {
            double gp = gin;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                gp = fma(-od[q], gp, fsc[q]);
                g[q] = gp;
            }
            if (pass + 1 < PARTS) {
                const double rg_ = from_prev_part<Gm::PARTD>(gp);
                gin = outermost ? 0.0 : rg_;
            }
        }
        const double xs = fma(-c_last, xchg<Gm::SEAM>(g[Q - 1]), g[Q - 1]) * r_seam;
        double xn = xs, x[Q];
#pragma unroll
        for (int pass = 0; pass < PARTS; ++pass) {
            x[Q - 1] = (pass == 0) ? xs : (innermost ? xs : fma(-c_last, xn, g[Q - 1]));
#pragma unroll
            for (int q = Q - 2; q >= 0; --q) x[q] = fma(-cp[q], x[q + 1], g[q]);
            if (pass + 1 < PARTS) xn = from_next_part<Gm::PARTD>(x[0]);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) U[q] -= x[q];
    }
    if (!col_ok) return;
    if (state_bc && top_lane) S.top_bc[c] = top;  // the cache keeps the flux of the last evaluation, as the reference's does
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int l = level_of(q);
        if (l < N) S.C[P.at(l, c)] = U[q];
    }
}

}  // namespace clb
