// soil_pair.cuh -- lane-pair variant of the fused implicit stage (9 <= N <= 16 levels).
//
// One column is owned by TWO lanes of a warp (lanes l and l + 16; 16 columns per warp):
// the "bottom" lane holds levels 0..7, the "top" lane levels N-1 down to 8, each as eight
// cell slots q = 0..7 ordered from the column's outer boundary towards the seam between
// levels 7 and 8.  In that orientation both halves run the SAME straight-line code:
//
//   * the closures (the FP64-heavy part) are evaluated four cells at a time with the four
//     dependency chains interleaved statement by statement (soil_mathv.cuh): the FP64 pipe
//     needs ~4 independent instructions in flight per sub-partition (DFMA: 8 cycles latency,
//     one warp-instruction per 2 cycles, tools/ubench/fp64_ilp.cu) and gets them from ILP,
//     not from occupancy;
//   * the stencil needs no shuffles inside a half; only the seam values cross lanes
//     (8 double shuffles per Newton iteration against ~150 of the lane-per-cell kernel);
//   * the tridiagonal systems are solved by the TWISTED (two-sided) Thomas factorisation:
//     each lane eliminates from its boundary towards the seam, the two seam rows give a
//     2x2 system solved redundantly by both lanes, and each lane back-substitutes outwards.
//     That is Thomas' operation count (~9 FP64 per cell against ~70 for 16-lane cyclic
//     reduction) at half its serial depth;
//   * per-cell stage constants (prepared closure parameters, lagged face coefficients, the
//     factored (rho_e, rho_e) block) live in SHARED memory, in a warp-private
//     [slot][level][column] tile (a lane only ever touches its own column: no barriers, and
//     a warp's access is two full 128-byte rows: no bank conflicts), so a column costs
//     ~1.8 KB of shared memory instead of 8 KB of registers; the iterate and the residual's
//     constant part stay in registers;
//   * every field is read from HBM once, by TMA: one cp.async.bulk.tensor per field and warp
//     brings the [N levels x 16 columns] box of the column-fastest mirror straight into that
//     tile (14 instructions per warp instead of 1800 per-lane loads), where it is transformed
//     in place into the stage constants; the new state is written once.
//
// Requires the column-fastest mirror layout (sl = ld, sc = 1).
//
// Inward flux convention: F_f is the flux through face f in the direction boundary -> seam
// (upward in the bottom half, downward in the top half), so for both halves
//     F_f = -a_f (h_inner - h_outer) [+ boundary flux],   T_q = (F_q - F_{q+1}) / dz_q
// which is the reference's -(q_hi - q_lo)/dz with exact sign flips (rre.jl:161-203,
// energy_hydrology.jl:363-425); Jacobian rows follow rre.jl:391-458 and
// energy_hydrology.jl:466-576; the block solve implicit_timestepping.jl:160-171.
#pragma once
#include <cuda.h>

#include "soil_device.cuh"
#include "soil_fused.cuh"
#include "soil_mathv.cuh"

namespace clb {

constexpr int kPairQ = 8;  // cell slots per lane
constexpr int kPairW = 4;  // cells evaluated together (interleaved dependency chains)

// Per-launch grid constants in the lane-local orientation, [half][slot]; a kernel
// parameter (constant bank), indexed with the lane's half at run time.
struct PairGrid {
    double z[2][kPairQ];          // z_c of the cell
    double dti[2][kPairQ];        // dtgamma / dz_c of the cell; 0 for pad slots
    double hidzf[2][kPairQ + 1];  // 1/(2 dz_f) of face f (f = q: outer face of slot q, 8: seam); 0 for boundary / pad faces
};

// Stage constants per cell: slots [0, NS) in shared memory, the rest in registers.
//   Richards (10)         theta_r, nu, ca, ca2, cb, 1/S_s, cc, cd, K_sat, t1
//   EnergyHydrology (17)  theta_r, nu_eff, ice energy, rho_c base, ca, K_lag rho_l c_l, cb, 1/S_s, cc, cd,
//                         aK_o, aC_o (outer-face coefficients), den22, od22 | c22, t1, t2
// t1 / t2: constant part of the residual, temp - dtgamma * (implicit source).
// Raw fields (what TMA brings in, same slots before the in-place transform):
//   Richards (9)          nu, theta_r, K_sat, S_s, a, b, m, theta_l, is_sat
//   EnergyHydrology (14)  nu, theta_r, S_s, a, b, m, theta_l, is_sat, theta_i, rho_c_ds, K, kappa, theta_l_lag, rho_e
template <int MODEL>
struct PairSlots {
    static constexpr int kConst = (MODEL == 1) ? 17 : 10;
    static constexpr int kRaw = (MODEL == 1) ? 14 : 9;
};
enum { R_THETA_R = 0, R_NU, R_CA, R_CA2, R_CB, R_INV_SS, R_CC, R_CD, R_KSAT, R_T1 };
enum { E_THETA_R = 0, E_NU_EFF, E_ICE, E_RCBASE, E_CA, E_KC, E_CB, E_INV_SS, E_CC, E_CD, E_AK, E_AC, E_DEN22, E_OD22,
       E_C22, E_T1, E_T2 };

constexpr int kPairTileBytes = 16 * 16 * 8;  // one slot of a warp: 16 level rows x 16 columns

// TMA descriptors of the raw fields (column-fastest mirrors: dims {ncol, N}, box {16, N})
struct PairMaps {
    CUtensorMap m[14];
};

template <int NS, int NTOT>
struct PairStore {
    double *base;  // warp tile + this lane's column
    int half;
    double r[(NTOT > NS) ? (NTOT - NS) : 1][kPairQ];
    // level row of cell slot q: bottom half q, top half 15 - q
    __device__ __forceinline__ double *at(int q, int slot) const
    {
        const int row = half ? 15 - q : q;
        return base + (slot * 16 + row) * 16;
    }
    template <int SLOT>
    __device__ __forceinline__ double get(int q) const
    {
        if constexpr (SLOT < NS)
            return *at(q, SLOT);
        else
            return r[SLOT - NS][q];
    }
    template <int SLOT>
    __device__ __forceinline__ void put(int q, double v)
    {
        if constexpr (SLOT < NS)
            *at(q, SLOT) = v;
        else
            r[SLOT - NS][q] = v;
    }
};

__device__ __forceinline__ double xchg(double v) { return __shfl_xor_sync(0xffffffffu, v, 16); }

// ---- TMA / mbarrier (sm_90+ PTX; one barrier per warp, single phase) -------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// box {16 columns, N levels} of a column-fastest mirror at (column c0, level 0) -> dense [level][16] tile
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int l0, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(l0), "r"(bar)
        : "memory");
}

// Prepared closure constants:
//   van Genuchten  ca = 1/m, ca2 = m, cb = 1/n, cc = 1/alpha, cd = 1/(alpha m n range)
//   Brooks-Corey   ca = -1/c, ca2 = 2/c + 3, cb = psi_b, cc = -psi_b/(c range), cd unused
struct ClosureConst {
    double ca, ca2, cb, cc, cd, inv_Ss;
};
template <int CLOSURE>
__device__ __forceinline__ ClosureConst pair_prepare(const HydroCell &p, double nu_eff)
{
    const double theta_lo = p.theta_r + kSqrtEps;
    const double range = fmax(nu_eff, theta_lo) - p.theta_r;
    ClosureConst c;
    c.inv_Ss = fm::rcp(p.S_s);
    if (CLOSURE == kVanGenuchten) {
        c.ca = fm::rcp(p.m);
        c.ca2 = p.m;
        c.cb = fm::rcp(p.b);
        c.cc = fm::rcp(p.a);
        c.cd = fm::rcp((p.a * p.m * p.b) * range);
    } else {
        c.ca = -fm::rcp(p.a);
        c.ca2 = fma(-2.0, c.ca, 3.0);
        c.cb = p.b;
        c.cc = -fm::div(p.b, p.a * range);
        c.cd = 0.0;
    }
    return c;
}

#ifndef CLB_PAIR_BLOCK
#define CLB_PAIR_BLOCK 64
#endif
#ifndef CLB_PAIR_MIN_BLOCKS
#define CLB_PAIR_MIN_BLOCKS 4
#endif

// Dynamic shared memory of a block: one NS-slot tile per warp [+ one mbarrier per warp when N == 16;
// for N < 16 the barrier sits in the (never loaded) level-15 row of slot 0 until the pad constants
// overwrite it].
template <int NS, int N, int BLOCK>
constexpr size_t pair_smem_bytes()
{
    return (size_t)(BLOCK / 32) * NS * kPairTileBytes + ((N == 16 || BLOCK != 64) ? (BLOCK / 32) * 8 : 0);
}

template <int CLOSURE, int MODEL, int N, int NS, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 64 ? CLB_PAIR_MIN_BLOCKS : 2))
    k_step_pair(const DevView P, const PairGrid G, const __grid_constant__ PairMaps M, double dtg, int max_iters)
{
    static_assert(N >= 9 && N <= 16, "lane-pair kernel: 9 <= N <= 16");
    constexpr int Q = kPairQ, W = kPairW;
    constexpr int Q0T = 16 - N;  // first real slot of the top half (pads before it)
    constexpr int NTOT = PairSlots<MODEL>::kConst, NRAW = PairSlots<MODEL>::kRaw;
    static_assert(NS >= NRAW && NS <= NTOT, "the raw fields are staged in the shared-memory slots");
    extern __shared__ __align__(128) unsigned char pair_sm[];

    const int tid = threadIdx.x, lane = tid & 31, half = lane >> 4, wib = tid >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * BLOCK + tid) >> 5;
    if (warp * 16 >= P.ncol) return;  // whole warp
    const int64_t c = warp * 16 + (lane & 15);
    const bool col_ok = c < P.ncol;
    const int64_t cs = col_ok ? c : P.ncol - 1;
    const EarthConst &E = P.earth;
    const double C1 = E.cp_l * E.rho_l, C2 = E.cp_i * E.rho_i, T_ref = E.T_ref;

    double *tile = reinterpret_cast<double *>(pair_sm + (size_t)wib * NS * kPairTileBytes);
    PairStore<NS, NTOT> S;
    S.base = tile + (lane & 15);
    S.half = half;

    // ---- every per-cell field of the stage: HBM -> shared memory, one TMA box per field --------
    const unsigned bar = (N == 16 || BLOCK != 64) ? smem_u32(pair_sm + (size_t)(BLOCK / 32) * NS * kPairTileBytes + wib * 8)
                                   : smem_u32(tile + 15 * 16);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (unsigned)(NRAW - ((CLOSURE == kVanGenuchten) ? 0 : 1)) * N * 16 * 8);
    }
    __syncwarp();
    {
        const bool skip = (CLOSURE != kVanGenuchten) && lane == ((MODEL == 1) ? 5 : 6);  // no m field for Brooks-Corey
        if (lane < NRAW && !skip)
            tma_load_2d(smem_u32(tile) + lane * kPairTileBytes, &M.m[lane], (int)(warp * 16), 0, bar);
    }

    // ---- per-column scalars (overlap the copies) ---------------------------------------------
    const double ld_Rss = P.R_ss[cs], ld_hg = P.h_grad[cs];
    const double top_w = P.top_bc_w[cs], bot_w = P.bot_bc_w[cs];
    double ld_Ress = 0.0, top_h = 0.0, bot_h = 0.0;
    if (MODEL == 1) {
        ld_Ress = P.R_ess[cs];
        top_h = P.top_bc_h[cs];
        bot_h = P.bot_bc_h[cs];
    }
    const double inv_hg = fm::rcp(fmax(ld_hg, kEps));
    const double src_w = ld_Rss * inv_hg, src_e = ld_Ress * inv_hg;
    // boundary flux of this half in the inward convention; it enters at face 0 (bottom half,
    // top half when N == 16) or at face Q0T (top half with pads)
    const double bin_w = half ? -top_w : bot_w, bin_e = half ? -top_h : bot_h;
    const double b0_w = (half && Q0T > 0) ? 0.0 : bin_w, b0_e = (half && Q0T > 0) ? 0.0 : bin_e;
    const double bT_w = half ? bin_w : 0.0, bT_e = half ? bin_e : 0.0;  // used at face Q0T > 0 only

    // flux integrals (W = -I, lagged boundary fluxes): their Newton recurrence does not depend on
    // the iterate, so it runs here, off the hot loop (one lane per column stores it)
    double dx2_int = 0.0;
    {
        const double tiw = P.Y_intF_w[cs];
        const double Tiw = -(top_w - bot_w) - ld_Rss;
        double Uw = tiw, dxw = 0.0, Ue = 0.0, dxe = 0.0, tie = 0.0, Tie = 0.0;
        if (MODEL == 1) {
            tie = P.Y_intF_e[cs];
            Tie = -(top_h - bot_h) - ld_Ress;
            Ue = tie;
        }
        for (int it = 0; it < max_iters; ++it) {
            dxw = -(tiw + dtg * Tiw - Uw);
            Uw -= dxw;
            dxe = -(tie + dtg * Tie - Ue);
            Ue -= dxe;
        }
        if (half == 0 && col_ok) {
            dx2_int = dxw * dxw + dxe * dxe;
            P.out_intF_w[c] = Uw;
            if (MODEL == 1) P.out_intF_e[c] = Ue;
        }
    }

    mbar_wait(bar, 0);
    __syncwarp();  // nobody polls the barrier any more
    if (N < 16 && BLOCK == 64) {  // its bytes are about to be overwritten by the pad constants
        if (lane == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        __syncwarp();
    }

    // ---- set-up: transform the raw fields in place into the stage constants ------------------
    double U1[Q], U2[Q];
    double aK8 = 0.0, aC8 = 0.0, r22 = 0.0, c22_7 = 0.0;
    if (MODEL == 0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const bool real = (half ? 15 - q : q) < N;
            HydroCell hc;
            hc.nu = S.template get<0>(q); hc.theta_r = S.template get<1>(q); hc.K_sat = S.template get<2>(q);
            hc.S_s = S.template get<3>(q); hc.a = S.template get<4>(q); hc.b = S.template get<5>(q);
            hc.m = S.template get<6>(q);
            double theta = S.template get<7>(q), sat = S.template get<8>(q);
            if (!real) {  // pad slot: benign parameters, identity rows (dti = 0, zero face coefficients)
                hc.nu = 0.5; hc.theta_r = 0.1; hc.K_sat = 0.0; hc.S_s = 1e-3; hc.a = 1.0; hc.b = 2.0; hc.m = 0.5;
                theta = 0.3; sat = 0.0;
            }
            const ClosureConst cc = pair_prepare<CLOSURE>(hc, hc.nu);
            U1[q] = theta;
            S.template put<R_THETA_R>(q, hc.theta_r);
            S.template put<R_NU>(q, hc.nu);
            S.template put<R_CA>(q, cc.ca);
            S.template put<R_CA2>(q, cc.ca2);
            S.template put<R_CB>(q, cc.cb);
            S.template put<R_INV_SS>(q, cc.inv_Ss);
            S.template put<R_CC>(q, cc.cc);
            S.template put<R_CD>(q, cc.cd);
            S.template put<R_KSAT>(q, hc.K_sat);
            S.template put<R_T1>(q, fma(-dtg, src_w * sat, theta));
        }
    } else {
        // rolling window over q-1, q, q+1 of the lagged fields that couple neighbours
        auto lagged = [&](int q, double &K, double &kap, double &rc) {
            const bool real = (half ? 15 - q : q) < N;
            K = real ? S.template get<10>(q) : 0.0;
            kap = real ? S.template get<11>(q) : 0.0;
            // Jacobian uses the LAGGED theta_l for rho_c_s (energy_hydrology.jl:561-566)
            rc = fm::rcp(volumetric_heat_capacity(real ? S.template get<12>(q) : 0.0, real ? S.template get<8>(q) : 0.0,
                                                  real ? S.template get<9>(q) : 1e6, E));
        };
        double K_m = 0.0, kap_m = 0.0, rc_m = 0.0, K_0, kap_0, rc_0, K_p = 0.0, kap_p = 0.0, rc_p = 0.0;
        lagged(0, K_0, kap_0, rc_0);
        double K7 = 0.0, kap7 = 0.0, rc7 = 0.0;
        lagged(Q - 1, K7, kap7, rc7);
        const double K7p = xchg(K7), kap7p = xchg(kap7), rc7p = xchg(rc7);
        aK8 = (K7 + K7p) * G.hidzf[half][Q];
        aC8 = (kap7 + kap7p) * G.hidzf[half][Q];
        double cprev = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const bool real = (half ? 15 - q : q) < N;
            if (q < Q - 1) lagged(q + 1, K_p, kap_p, rc_p);
            HydroCell hc;
            hc.nu = S.template get<0>(q); hc.theta_r = S.template get<1>(q); hc.K_sat = 0.0;
            hc.S_s = S.template get<2>(q); hc.a = S.template get<3>(q); hc.b = S.template get<4>(q);
            hc.m = S.template get<5>(q);
            double theta = S.template get<6>(q), sat = S.template get<7>(q), theta_i = S.template get<8>(q),
                   rcds = S.template get<9>(q), rho_e = S.template get<13>(q);
            if (!real) {
                hc.nu = 0.5; hc.theta_r = 0.1; hc.S_s = 1e-3; hc.a = 1.0; hc.b = 2.0; hc.m = 0.5;
                theta = 0.3; sat = 0.0; theta_i = 0.0; rcds = 1e6; rho_e = 0.0;
            }
            const double nu_eff = hc.nu - theta_i;
            const ClosureConst cc = pair_prepare<CLOSURE>(hc, nu_eff);
            U1[q] = theta;
            U2[q] = rho_e;
            // lagged face coefficients and the row of W22 = dtgamma d(T_rho_e)/d(rho_e) - I, eliminated on the fly
            const double hid_o = G.hidzf[half][q], dti = G.dti[half][q];
            const double aK_o = (q == 0) ? 0.0 : (K_0 + K_m) * hid_o;
            const double aC_o = (q == 0) ? 0.0 : (kap_0 + kap_m) * hid_o;
            const double aC_in = (q < Q - 1) ? (kap_0 + kap_p) * G.hidzf[half][q + 1] : aC8;
            const double o = (q == 0) ? 0.0 : (aC_o * rc_m) * dti;
            const double i = (aC_in * ((q < Q - 1) ? rc_p : rc7p)) * dti;
            const double d = fma(-((aC_in + aC_o) * rc_0), dti, -1.0);
            const double den = fm::rcp(fma(-o, cprev, d));
            cprev = i * den;
            S.template put<E_THETA_R>(q, hc.theta_r);
            S.template put<E_NU_EFF>(q, nu_eff);
            S.template put<E_ICE>(q, theta_i * E.rho_i * E.LH_f0);
            S.template put<E_RCBASE>(q, fma(theta_i, C2, rcds));
            S.template put<E_CA>(q, cc.ca);
            S.template put<E_KC>(q, K_0 * C1);
            S.template put<E_CB>(q, cc.cb);
            S.template put<E_INV_SS>(q, cc.inv_Ss);
            S.template put<E_CC>(q, cc.cc);
            S.template put<E_CD>(q, cc.cd);
            S.template put<E_AK>(q, aK_o);
            S.template put<E_AC>(q, aC_o);
            S.template put<E_DEN22>(q, den);
            S.template put<E_OD22>(q, o * den);
            S.template put<E_C22>(q, cprev);
            S.template put<E_T1>(q, fma(-dtg, src_w * sat, theta));
            S.template put<E_T2>(q, fma(-dtg, src_e * sat, rho_e));
            K_m = K_0; kap_m = kap_0; rc_m = rc_0;
            K_0 = K_p; kap_0 = kap_p; rc_0 = rc_p;
        }
        c22_7 = cprev;
        r22 = fm::rcp(fma(-c22_7, xchg(c22_7), 1.0));
    }

    // ---- Newton iterations -----------------------------------------------------------------
    double dx2 = 0.0;
#pragma unroll 1
    for (int it = 0; it < max_iters; ++it) {
        // cache_imp!: closures (and temperature) at the iterate, four cells at a time
        double h[Q], dps[Q], Kc[Q], Td[Q], eK[Q];
#pragma unroll
        for (int g = 0; g < Q; g += W) {
            double th[W], thr[W], nue[W], ca[W], ca2[W], cb[W], iSs[W], ccc[W], cd[W], Ksat[W];
            double K[W], psi[W], dp[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                th[j] = U1[g + j];
                thr[j] = S.template get<0>(g + j);  // theta_r / nu(_eff) are slots 0 / 1 of both models
                nue[j] = S.template get<1>(g + j);
                ca2[j] = 0.0;
                Ksat[j] = 0.0;
            }
            if (MODEL == 1) {
                // update_implicit_aux (energy_hydrology.jl:427-445): T from (theta_l clipped to the pore
                // space left by ice, rho_e_int, theta_i)
                double num[W], rcs[W], Tq[W];
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    num[j] = U2[g + j] + S.template get<E_ICE>(g + j);
                    rcs[j] = fma(fmv::min_nn(nue[j], th[j]), C1, S.template get<E_RCBASE>(g + j));
                }
                fmv::div<W>(num, rcs, Tq);
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const double T = T_ref + Tq[j];
                    Td[g + j] = T;
                    eK[g + j] = (T - T_ref) * S.template get<E_KC>(g + j);
                }
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    ca[j] = S.template get<E_CA>(g + j);
                    cb[j] = S.template get<E_CB>(g + j);
                    iSs[j] = S.template get<E_INV_SS>(g + j);
                    ccc[j] = S.template get<E_CC>(g + j);
                    cd[j] = S.template get<E_CD>(g + j);
                }
            } else {
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    ca[j] = S.template get<R_CA>(g + j);
                    ca2[j] = S.template get<R_CA2>(g + j);
                    cb[j] = S.template get<R_CB>(g + j);
                    iSs[j] = S.template get<R_INV_SS>(g + j);
                    ccc[j] = S.template get<R_CC>(g + j);
                    cd[j] = S.template get<R_CD>(g + j);
                    Ksat[j] = S.template get<R_KSAT>(g + j);
                }
            }
            fmv::closure<CLOSURE, MODEL == 0, W>(th, thr, nue, ca, ca2, cb, ccc, cd, iSs, Ksat, K, psi, dp);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                h[g + j] = psi[j] + G.z[half][g + j];
                dps[g + j] = dp[j];
                if (MODEL == 0) Kc[g + j] = K[j];
            }
        }
        const double h7p = xchg(h[Q - 1]), dps7p = xchg(dps[Q - 1]);
        double K7p = 0.0, T7p = 0.0, eK7p = 0.0;
        if (MODEL == 0) K7p = xchg(Kc[Q - 1]);
        if (MODEL == 1) {
            T7p = xchg(Td[Q - 1]);
            eK7p = xchg(eK[Q - 1]);
        }

        // T_imp!, Wfact and the forward elimination of W11, boundary -> seam
        double c1[Q], g1[Q], f2[Q], aE[Q + 1];
        double Fw_o = b0_w, Fe_o = b0_e, aK_o = 0.0;
        double cprev = 0.0, gprev = 0.0;
        aE[0] = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double dti = G.dti[half][q];
            const double hid_in = G.hidzf[half][q + 1];
            const double h_in = (q < Q - 1) ? h[q + 1] : h7p;
            const double dps_in = (q < Q - 1) ? dps[q + 1] : dps7p;
            const double dh = h_in - h[q];
            double aK_in, aC_in = 0.0;  // coefficients of the inner face
            if (MODEL == 1) {
                aK_in = (q < Q - 1) ? S.template get<E_AK>(q + 1) : aK8;
                aC_in = (q < Q - 1) ? S.template get<E_AC>(q + 1) : aC8;
            } else {
                aK_in = (Kc[q] + ((q < Q - 1) ? Kc[q + 1] : K7p)) * hid_in;
            }
            double Fw_in = -(aK_in * dh);
            if (Q0T > 0 && q + 1 == Q0T) Fw_in += bT_w;
            double t1;
            if (MODEL == 0) t1 = S.template get<R_T1>(q);
            else t1 = S.template get<E_T1>(q);
            const double f1 = fma(Fw_o - Fw_in, dti, t1) - U1[q];
            // row of W11 = dtgamma dT/dtheta - I in the lane-local orientation
            const double o = (q == 0) ? 0.0 : (aK_o * dps[q - 1]) * dti;
            const double i = (aK_in * dps_in) * dti;
            const double d = fma(-((aK_in + aK_o) * dps[q]), dti, -1.0);
            const double den = fm::rcp(fma(-o, cprev, d));
            cprev = i * den;
            gprev = fma(-o, gprev, f1) * den;
            c1[q] = cprev;
            g1[q] = gprev;
            if (MODEL == 1) {
                const double T_in = (q < Q - 1) ? Td[q + 1] : T7p;
                const double eK_in = (q < Q - 1) ? eK[q + 1] : eK7p;
                const double aE_in = (eK[q] + eK_in) * hid_in;
                aE[q + 1] = aE_in;
                double Fe_in = fma(-aC_in, T_in - Td[q], -(aE_in * dh));
                if (Q0T > 0 && q + 1 == Q0T) Fe_in += bT_e;
                f2[q] = fma(Fe_o - Fe_in, dti, S.template get<E_T2>(q)) - U2[q];
                Fe_o = Fe_in;
            }
            Fw_o = Fw_in;
            aK_o = aK_in;
        }
        // seam: x_7 + c_7 x'_7 = g_7 in both lanes
        const double c7p = xchg(c1[Q - 1]), g7p = xchg(g1[Q - 1]);
        double x1[Q], y[Q];
        x1[Q - 1] = fma(-c1[Q - 1], g7p, g1[Q - 1]) * fm::rcp(fma(-c1[Q - 1], c7p, 1.0));
#pragma unroll
        for (int q = Q - 2; q >= 0; --q) x1[q] = fma(-c1[q], x1[q + 1], g1[q]);
        const bool last = (it == max_iters - 1);
        if (last) {
            dx2 = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) dx2 = fma(x1[q], x1[q], dx2);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            U1[q] -= x1[q];
            y[q] = dps[q] * x1[q];
        }
        if (MODEL == 1) {
            // ldiv!: BlockLowerTriangularSolve(theta_l): b2 = f2 - W21 x1 with
            // W21 = -dtgamma (D . Diag(interp(-e_l K)) . G . Diag(dpsi)) - I  (energy_hydrology.jl:545-556)
            const double y7p = xchg(y[Q - 1]);
            double g2[Q];
            double g2prev = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double dti = G.dti[half][q];
                const double y_in = (q < Q - 1) ? y[q + 1] : y7p;
                double acc = aE[q + 1] * (y_in - y[q]);
                if (q > 0) acc = fma(aE[q], y[q - 1] - y[q], acc);
                const double s = fma(dti, acc, -x1[q]);
                const double b2 = f2[q] - s;
                g2prev = fma(-S.template get<E_OD22>(q), g2prev, b2 * S.template get<E_DEN22>(q));
                g2[q] = g2prev;
            }
            const double g27p = xchg(g2[Q - 1]);
            double x2 = fma(-c22_7, g27p, g2[Q - 1]) * r22;
            U2[Q - 1] -= x2;
            if (last) dx2 = fma(x2, x2, dx2);
#pragma unroll
            for (int q = Q - 2; q >= 0; --q) {
                x2 = fma(-S.template get<E_C22>(q), x2, g2[q]);
                U2[q] -= x2;
                if (last) dx2 = fma(x2, x2, dx2);
            }
        }
    }

    // ---- write the new state -----------------------------------------------------------------
    double bad = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int level = half ? 15 - q : q;
        if (level < N && col_ok) {
            const int64_t k = P.at(level, c);
            P.out_theta_l[k] = U1[q];
            if (!isfinite(U1[q])) bad += 1.0;
            if (MODEL == 1) {
                P.out_rho_e[k] = U2[q];
                if (!isfinite(U2[q])) bad += 1.0;
            }
        }
    }
    if (!col_ok) dx2 = 0.0;
    accumulate_stats(P, dx2 + dx2_int, bad);
}

}  // namespace clb
