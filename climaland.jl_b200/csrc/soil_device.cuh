// soil_device.cuh -- device-side view of one handle's fields (column-fastest SoA).
#pragma once
#include <stdint.h>

#include "soil_closures.cuh"

namespace clb {

constexpr int kMaxStaticLevels = 64;  // grid vectors of the register-column kernels travel as kernel parameters

// Everything a kernel reads or writes.  Per-cell arrays: element (level i,
// column c) at base[i*sl + c*sc]; per-column arrays: base[c].  Two mirror layouts:
//   column-fastest  sl = ld (ncol rounded up to 32), sc = 1   -- thread-per-column kernels
//   level-fastest   sl = 1, sc = N (the reference's own layout) -- lane-per-cell kernels
// Pointers of fields the configured model does not use are null.
struct DevView {
    int32_t model, closure, top_bc, bottom_bc, topmodel;
    int32_t N;
    int64_t ncol, ld;
    int64_t sl, sc;
    __device__ __forceinline__ int64_t at(int i, int64_t c) const { return (int64_t)i * sl + c * sc; }
    EarthConst earth;
    // grid (device arrays, length N; inv_dz_f[i] belongs to the face between cells i-1 and i, i = 1..N-1)
    const double *z_c, *dz_c, *inv_dz_c, *inv_dz_f;
    double dz_top, dz_bot;  // Domains.get_dz top / bottom half-cell, Domains.jl:935-949
    double inv_dz_top, inv_dz_bot;  // 1 / dz_top, 1 / dz_bot divided on the host (correctly rounded): fm::div_by
    // parameters and lagged cache
    const double *nu, *theta_r, *K_sat, *S_s, *hcm_a, *hcm_b, *hcm_m, *rho_c_ds;
    const double *K_lag, *kappa_lag, *theta_l_lag, *is_sat;
    const double *R_ss, *R_ess, *h_grad, *theta_bc_top, *theta_bc_bot;
    // state, cache, tendency
    double *Y_theta_l, *Y_rho_e, *Y_theta_i, *Y_intF_w, *Y_intF_e;
    // where the fused stage writes the new state: the Y fields (in place) or the U fields (out of place)
    double *out_theta_l, *out_rho_e, *out_intF_w, *out_intF_e;
    double *p_K, *p_psi, *p_T;
    double *top_bc_w, *bot_bc_w, *top_bc_h, *bot_bc_h, *dfluxBCdY, *total_water;
    double *dY_theta_l, *dY_rho_e, *dY_theta_i, *dY_intF_w, *dY_intF_e;
    // Jacobian
    double *w11_lo, *w11_di, *w11_up, *w21_lo, *w21_di, *w21_up, *w22_lo, *w22_di, *w22_up;
    // linear solve
    const double *b_theta_l, *b_rho_e, *b_theta_i, *b_intF_w, *b_intF_e;
    double *x_theta_l, *x_rho_e, *x_theta_i, *x_intF_w, *x_intF_e;
    // scratch: cell-sized arrays (iterate and Thomas vectors of the generic kernels, ldiv work)
    double *work[6];
    // per-column scalars carried between the launches of the tolerance path (4 rows of ld)
    double *carry;
    // step statistics: [0] = sum dx^2 (last iteration), [1] = non-finite count (as double)
    double *stats;
    // convergence flag of the tolerance path (device int; 1 = converged, later iterations are skipped)
    int32_t *converged;
};

__device__ __forceinline__ HydroCell load_cell(const DevView &P, int64_t k)
{
    HydroCell h;
    h.nu = __ldg(P.nu + k);
    h.theta_r = __ldg(P.theta_r + k);
    h.K_sat = __ldg(P.K_sat + k);
    h.S_s = __ldg(P.S_s + k);
    h.a = __ldg(P.hcm_a + k);
    h.b = __ldg(P.hcm_b + k);
    h.m = P.hcm_m ? __ldg(P.hcm_m + k) : 0.0;
    return h;
}

// Grid vectors passed by value to the register-column kernels: uniform, compile-time
// indexed, so they become constant-bank operands of the FP64 instructions.
template <int N>
struct GridConst {
    double z_c[N];
    double inv_dz_c[N];
    double inv_dz_f[N];  // [i]: face between i-1 and i (i >= 1); [0] unused
};

}  // namespace clb
