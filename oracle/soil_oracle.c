/*
 * soil_oracle.c -- CPU restatement of ClimaLand.jl's implicit soil-column path.
 * TEST INFRASTRUCTURE ONLY (see soil_oracle.h).  Every function cites the
 * reference lines it follows; paths are relative to /root/reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  FP
 * contraction is off so a*b+c rounds twice, as Julia does without muladd.
 */
#include "soil_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SQRT_EPS 1.4901161193847656e-8 /* sqrt(eps(Float64)) */
#define EPS64 DBL_EPSILON              /* eps(Float64) = 2.220446049250313e-16 */

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ */
/* point functions                                                     */
/* ------------------------------------------------------------------ */

/* src/standalone/Soil/soil_hydrology_parameterizations.jl:45-50 */
double orc_effective_saturation(double nu_eff, double theta_l, double theta_r)
{
    double theta_safe = dmax(theta_l, theta_r + SQRT_EPS);
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    return (theta_safe - theta_r) / (nu_safe - theta_r);
}

/* soil_hydrology_parameterizations.jl:22-31 */
double orc_volumetric_liquid_fraction(double theta_l, double nu_eff, double theta_r)
{
    double theta_safe = dmax(theta_l, theta_r + SQRT_EPS);
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    return theta_safe < nu_safe ? theta_safe : nu_safe;
}

/* soil_hydrology_parameterizations.jl:59-63 */
double orc_vg_matric_potential(double alpha, double n, double m, double S)
{
    return -pow((pow(S, -1.0 / m) - 1.0) * pow(alpha, -n), 1.0 / n);
}

/* soil_hydrology_parameterizations.jl:72-77 (psi > 0 is an error upstream; NaN here) */
double orc_vg_inverse_matric_potential(double alpha, double n, double m, double psi)
{
    if (psi > 0) return NAN;
    return pow(1.0 + pow(alpha * fabs(psi), n), -m);
}

/* soil_hydrology_parameterizations.jl:109-127 */
double orc_vg_pressure_head(double alpha, double n, double m, double theta_r,
                            double theta_l, double nu_eff, double S_s)
{
    double S = orc_effective_saturation(nu_eff, theta_l, theta_r);
    double theta_safe = dmax(theta_l, theta_r + SQRT_EPS);
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    if (S <= 1.0) return orc_vg_matric_potential(alpha, n, m, S);
    return (theta_safe - nu_safe) / S_s;
}

/* soil_hydrology_parameterizations.jl:135-152 */
double orc_vg_dpsidtheta(double alpha, double n, double m, double theta,
                         double nu_eff, double theta_r, double S_s)
{
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    double S = orc_effective_saturation(nu_safe, theta, theta_r);
    if (S < 1.0) {
        return 1.0 / (alpha * m * n) / (nu_safe - theta_r) *
               pow(pow(S, -1.0 / m) - 1.0, 1.0 / n - 1.0) * pow(S, -1.0 / m - 1.0);
    }
    return 1.0 / S_s;
}

/* soil_hydrology_parameterizations.jl:161-173 */
double orc_vg_hydraulic_conductivity(double m, double K_sat, double S)
{
    double K;
    if (S < 1.0) {
        double t = 1.0 - pow(1.0 - pow(S, 1.0 / m), m);
        K = sqrt(S) * (t * t); /* x^FT(2) lowers to x*x */
    } else {
        K = 1.0;
    }
    return K * K_sat;
}

/* soil_hydrology_parameterizations.jl:182-186 */
double orc_bc_matric_potential(double c, double psi_b, double S)
{
    return psi_b * pow(S, -1.0 / c);
}

/* soil_hydrology_parameterizations.jl:195-200 */
double orc_bc_inverse_matric_potential(double c, double psi_b, double psi)
{
    if (psi > 0) return NAN;
    return pow(psi / psi_b, -c);
}

/* soil_hydrology_parameterizations.jl:270-289 */
double orc_bc_pressure_head(double c, double psi_b, double theta_r, double theta_l,
                            double nu_eff, double S_s)
{
    double S = orc_effective_saturation(nu_eff, theta_l, theta_r);
    double theta_safe = dmax(theta_l, theta_r + SQRT_EPS);
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    if (S <= 1.0) return orc_bc_matric_potential(c, psi_b, S);
    return (theta_safe - nu_safe) / S_s + psi_b;
}

/* soil_hydrology_parameterizations.jl:220-231 */
double orc_bc_dpsidtheta(double c, double psi_b, double theta, double nu_eff,
                         double theta_r, double S_s)
{
    double nu_safe = dmax(nu_eff, theta_r + SQRT_EPS);
    double S = orc_effective_saturation(nu_safe, theta, theta_r);
    if (S < 1.0) return -psi_b / (c * (nu_safe - theta_r)) * pow(S, -(1.0 + 1.0 / c));
    return 1.0 / S_s;
}

/* soil_hydrology_parameterizations.jl:239-251 */
double orc_bc_hydraulic_conductivity(double c, double K_sat, double S)
{
    double K = (S < 1.0) ? pow(S, 2.0 / c + 3.0) : 1.0;
    return K * K_sat;
}

/* soil_hydrology_parameterizations.jl:302-305 */
double orc_impedance_factor(double f_i, double Omega) { return pow(10.0, -Omega * f_i); }

/* soil_hydrology_parameterizations.jl:320-324 */
double orc_viscosity_factor(double T, double gamma, double gammaT_ref)
{
    return exp(gamma * (T - gammaT_ref));
}

/* src/standalone/Soil/soil_heat_parameterizations.jl:157-171 */
double orc_volumetric_heat_capacity(double theta_l, double theta_i, double rho_c_ds,
                                    double rho_l, double cp_l, double rho_i, double cp_i)
{
    double rhocp_i = cp_i * rho_i;
    double rhocp_l = cp_l * rho_l;
    return rho_c_ds + theta_l * rhocp_l + theta_i * rhocp_i;
}

/* soil_heat_parameterizations.jl:180-192 */
double orc_temperature_from_rho_e_int(double rho_e_int, double theta_i, double rho_c_s,
                                      double rho_i, double T_ref, double LH_f0)
{
    return T_ref + (rho_e_int + theta_i * rho_i * LH_f0) / rho_c_s;
}

/* soil_heat_parameterizations.jl:201-212 */
double orc_volumetric_internal_energy(double theta_i, double rho_c_s, double T,
                                      double rho_i, double T_ref, double LH_f0)
{
    return rho_c_s * (T - T_ref) - theta_i * rho_i * LH_f0;
}

/* soil_heat_parameterizations.jl:221-231 */
double orc_volumetric_internal_energy_liq(double T, double rho_l, double cp_l, double T_ref)
{
    double rhocp_l = cp_l * rho_l;
    return rhocp_l * (T - T_ref);
}

/* src/shared_utilities/utils.jl:93-99 */
double orc_heaviside(double x, double a) { return (x - a > EPS64) ? 1.0 : 0.0; }

/* ------------------------------------------------------------------ */
/* closure dispatch                                                    */
/* ------------------------------------------------------------------ */
static inline double hcm_pressure_head(const orc_problem *P, int64_t k, double theta_l, double nu_eff)
{
    if (P->closure == ORC_VAN_GENUCHTEN)
        return orc_vg_pressure_head(P->hcm_a[k], P->hcm_b[k], P->hcm_m[k], P->theta_r[k],
                                    theta_l, nu_eff, P->S_s[k]);
    return orc_bc_pressure_head(P->hcm_a[k], P->hcm_b[k], P->theta_r[k], theta_l, nu_eff, P->S_s[k]);
}

static inline double hcm_dpsidtheta(const orc_problem *P, int64_t k, double theta_l, double nu_eff)
{
    if (P->closure == ORC_VAN_GENUCHTEN)
        return orc_vg_dpsidtheta(P->hcm_a[k], P->hcm_b[k], P->hcm_m[k], theta_l, nu_eff,
                                 P->theta_r[k], P->S_s[k]);
    return orc_bc_dpsidtheta(P->hcm_a[k], P->hcm_b[k], theta_l, nu_eff, P->theta_r[k], P->S_s[k]);
}

static inline double hcm_conductivity(const orc_problem *P, int64_t k, double S)
{
    if (P->closure == ORC_VAN_GENUCHTEN)
        return orc_vg_hydraulic_conductivity(P->hcm_m[k], P->K_sat[k], S);
    return orc_bc_hydraulic_conductivity(P->hcm_a[k], P->K_sat[k], S);
}

/* Domains.get_dz: src/shared_utilities/Domains.jl:935-949 (top/bottom = half the
 * end cell; pinned by test/standalone/Soil/soil_bc.jl:110-115) */
static inline double dz_cell(const orc_problem *P, int i) { return P->z_f[i + 1] - P->z_f[i]; }
static inline double dz_top(const orc_problem *P) { return dz_cell(P, P->N - 1) / 2.0; }
static inline double dz_bottom(const orc_problem *P) { return dz_cell(P, 0) / 2.0; }
/* distance between the centres either side of interior face f (1..N-1) */
static inline double dz_face(const orc_problem *P, int f) { return P->z_c[f] - P->z_c[f - 1]; }

#define FOR_COLUMNS(P, c)                                                                  \
    _Pragma("omp parallel for schedule(static) if ((P)->nthreads > 1) num_threads((P)->nthreads > 1 ? (P)->nthreads : 1)") \
    for (int64_t c = 0; c < (P)->ncol; ++c)

/* ------------------------------------------------------------------ */
/* boundary fluxes                                                     */
/* ------------------------------------------------------------------ */

/* One column of update_boundary_fluxes! for the water equation:
 *   rre.jl:111-149 -> boundary_flux! methods,
 *   src/standalone/Soil/boundary_conditions.jl:227-267 (MoistureStateBC top),
 *   :282-325 (MoistureStateBC bottom), :340-353 (FreeDrainage),
 *   set_dfluxBCdY! :375-411, diffusive_flux
 *   src/shared_utilities/boundary_conditions.jl:63-65.
 * Flux-type BCs (WaterFluxBC, atmos-driven) are values supplied by the host. */
static void column_water_boundary_fluxes(const orc_problem *P, const orc_state *Y, orc_cache *p,
                                         int64_t c, const double *Kcol)
{
    const int N = P->N;
    const int64_t o = c * N;
    if (P->top_bc == ORC_TOP_MOISTURE_STATE) {
        int64_t k = o + N - 1;
        double dz = dz_top(P);
        /* boundary_flux! passes nu, not nu - theta_i, for either model */
        double psi_bc = hcm_pressure_head(P, k, P->theta_bc_top[c], P->nu[k]);
        /* diffusive_flux(K_eff, psi_bc + dz, psi_c, dz) = -K (x2 - x1)/dz */
        p->top_bc_w[c] = -Kcol[N - 1] * ((psi_bc + dz) - p->psi[k]) / dz;
        if (P->model == ORC_RICHARDS) {
            /* covariant3_unit_vector(...) * (K_N * dpsidtheta / dz); the unit-vector norm
             * is carried by the divergence below, so the scalar stored is K dpsi/dz. */
            p->dfluxBCdY[c] = Kcol[N - 1] * hcm_dpsidtheta(P, k, Y->theta_l[k], P->nu[k]) / dz;
        }
    }
    if (P->bottom_bc == ORC_BOT_FREE_DRAINAGE) {
        p->bot_bc_w[c] = -1 * Kcol[0];
    } else if (P->bottom_bc == ORC_BOT_MOISTURE_STATE) {
        int64_t k = o;
        double dz = dz_bottom(P);
        double psi_bc = hcm_pressure_head(P, k, P->theta_bc_bot[c], P->nu[k]);
        p->bot_bc_w[c] = -Kcol[0] * ((p->psi[k] + dz) - psi_bc) / dz;
    }
}

void orc_update_boundary_fluxes(const orc_problem *P, const orc_state *Y, orc_cache *p)
{
    FOR_COLUMNS(P, c)
    {
        const double *Kcol = (P->model == ORC_ENERGY_HYDROLOGY ? P->K_lag : p->K) + c * P->N;
        column_water_boundary_fluxes(P, Y, p, c, Kcol);
    }
}

/* ------------------------------------------------------------------ */
/* update_implicit_cache!                                              */
/* ------------------------------------------------------------------ */

/* src/shared_utilities/models.jl:238-246: update_implicit_aux then
 * update_implicit_boundary_fluxes.
 *  Richards: implicit aux == update_aux! (models.jl:207-210, rre.jl:368-380):
 *     K, psi, total_water; boundary fluxes re-evaluated only when dfluxBCdY is
 *     in the cache, i.e. top BC is MoistureStateBC (rre.jl:460-468).
 *  EnergyHydrology: T and psi only (energy_hydrology.jl:427-445); K, kappa,
 *     theta_l stay lagged; its cache never holds dfluxBCdY (boundary_vars default,
 *     src/shared_utilities/boundary_conditions.jl:102) so BCs stay lagged too. */
void orc_update_implicit_cache(const orc_problem *P, const orc_state *Y, orc_cache *p)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        if (P->model == ORC_RICHARDS) {
            double tw = 0.0;
            for (int i = 0; i < N; ++i) {
                int64_t k = o + i;
                double S = orc_effective_saturation(P->nu[k], Y->theta_l[k], P->theta_r[k]);
                p->K[k] = hcm_conductivity(P, k, S);
                p->psi[k] = hcm_pressure_head(P, k, Y->theta_l[k], P->nu[k]);
                tw += Y->theta_l[k] * dz_cell(P, i);
            }
            if (p->total_water) p->total_water[c] = tw;
            if (P->top_bc == ORC_TOP_MOISTURE_STATE)
                column_water_boundary_fluxes(P, Y, p, c, p->K + o);
        } else {
            for (int i = 0; i < N; ++i) {
                int64_t k = o + i;
                double theta_i = Y->theta_i[k];
                double theta_l = dmin(P->nu[k] - theta_i, Y->theta_l[k]);
                double rho_c_s = orc_volumetric_heat_capacity(theta_l, theta_i, P->rho_c_ds[k],
                                                              P->rho_l, P->cp_l, P->rho_i, P->cp_i);
                p->T[k] = orc_temperature_from_rho_e_int(Y->rho_e_int[k], theta_i, rho_c_s,
                                                         P->rho_i, P->T_ref, P->LH_f0);
                p->psi[k] = hcm_pressure_head(P, k, Y->theta_l[k], P->nu[k] - theta_i);
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* compute_imp_tendency!                                               */
/* ------------------------------------------------------------------ */

/* rre.jl:161-203 and energy_hydrology.jl:363-425.
 * Operators (ClimaCore, un-vendored): InterpolateC2F = arithmetic mean of the two
 * adjacent centres, GradientC2F = centre difference over the centre spacing,
 * DivergenceF2C(SetValue top/bottom) = face-flux difference over the cell
 * thickness with the boundary face fluxes set to top_bc / bottom_bc.  Pinned by
 * test/standalone/Soil/soiltest.jl:357-406 (hand-built face-flux formula),
 * :82-86 (hydrostatic => 0), conservation.jl:139-151 (sign of boundary terms).
 * Implicit source: Runoff/Runoff.jl:321-359. */
void orc_compute_imp_tendency(const orc_problem *P, const orc_state *Y, const orc_cache *p,
                              orc_state *dY)
{
    (void)Y;
    const int N = P->N;
    const int eh = (P->model == ORC_ENERGY_HYDROLOGY);
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        const double *K = (eh ? P->K_lag : p->K) + o;
        const double *psi = p->psi + o;
        const double top_w = p->top_bc_w[c], bot_w = p->bot_bc_w[c];
        dY->intF_w[c] = -(top_w - bot_w);
        double top_h = 0, bot_h = 0;
        if (eh) {
            top_h = p->top_bc_h[c];
            bot_h = p->bot_bc_h[c];
            dY->intF_e[c] = -(top_h - bot_h);
        }
        /* sweep faces bottom -> top, carrying the flux through the lower face */
        double qw_lo = bot_w, qe_lo = bot_h;
        for (int i = 0; i < N; ++i) {
            double qw_hi, qe_hi = 0;
            if (i < N - 1) {
                double dzf = dz_face(P, i + 1);
                double grad_h = ((psi[i + 1] + P->z_c[i + 1]) - (psi[i] + P->z_c[i])) / dzf;
                qw_hi = -((K[i] + K[i + 1]) / 2.0) * grad_h;
                if (eh) {
                    const double *T = p->T + o, *kap = P->kappa_lag + o;
                    double e0 = orc_volumetric_internal_energy_liq(T[i], P->rho_l, P->cp_l, P->T_ref) * K[i];
                    double e1 = orc_volumetric_internal_energy_liq(T[i + 1], P->rho_l, P->cp_l, P->T_ref) * K[i + 1];
                    double grad_T = (T[i + 1] - T[i]) / dzf;
                    qe_hi = -((kap[i] + kap[i + 1]) / 2.0) * grad_T - ((e0 + e1) / 2.0) * grad_h;
                }
            } else {
                qw_hi = top_w;
                qe_hi = top_h;
            }
            double dzc = dz_cell(P, i);
            dY->theta_l[o + i] = -((qw_hi - qw_lo) / dzc);
            if (eh) {
                dY->rho_e_int[o + i] = -((qe_hi - qe_lo) / dzc);
                dY->theta_i[o + i] = 0.0;
            }
            qw_lo = qw_hi;
            qe_lo = qe_hi;
        }
        if (P->has_topmodel_source) {
            double hg = dmax(P->h_grad[c], EPS64);
            for (int i = 0; i < N; ++i) {
                dY->theta_l[o + i] -= (P->R_ss[c] / hg) * P->is_saturated[o + i];
                if (eh) dY->rho_e_int[o + i] -= (P->R_ess[c] / hg) * P->is_saturated[o + i];
            }
            dY->intF_w[c] -= P->R_ss[c];
            if (eh) dY->intF_e[c] -= P->R_ess[c];
        }
    }
}

/* ------------------------------------------------------------------ */
/* compute_jacobian!                                                   */
/* ------------------------------------------------------------------ */

/* One tridiagonal block  W = -dtgamma * (D . Diag(interp(-A)) . G . Diag(coef)) - I
 * for centre field A (K, e_liq*K or kappa) and centre coefficient coef (dpsi/dtheta or
 * 1/rho_c_s).  G has zero rows at the two boundary faces (SetGradient(0)).
 * rre.jl:423-454; energy_hydrology.jl:503-573.  Entries pinned by
 * test/shared_utilities/implicit_timestepping/richards_model.jl:92-141,192-225 and
 * energy_hydrology_model.jl:89-172. */
static void column_tridiag_block(const orc_problem *P, const double *A, const double *coef,
                                 double dtgamma, double top_dflux, double *lo, double *di, double *up)
{
    const int N = P->N;
    const double neg_dtg = -dtgamma;
    for (int i = 0; i < N; ++i) {
        double dzc = dz_cell(P, i);
        /* face below (i-1/2) and above (i+1/2): -interp(A)/dz_f, zero at the boundaries */
        double f_lo = 0.0, f_hi = 0.0;
        if (i > 0) f_lo = -((A[i - 1] + A[i]) / 2.0) / dz_face(P, i);
        if (i < N - 1) f_hi = -((A[i] + A[i + 1]) / 2.0) / dz_face(P, i + 1);
        /* dF_{i+1/2}/dY_i = -f_hi*coef_i, dF_{i+1/2}/dY_{i+1} = f_hi*coef_{i+1}; same below */
        double dFhi_dYi = -f_hi * coef[i];
        double dFhi_dYp = (i < N - 1) ? f_hi * coef[i + 1] : 0.0;
        double dFlo_dYm = (i > 0) ? -f_lo * coef[i - 1] : 0.0;
        double dFlo_dYi = f_lo * coef[i];
        if (i == N - 1) dFhi_dYi += top_dflux; /* LowerDiagonalMatrixRow(topBC_scratch), rre.jl:434-449 */
        lo[i] = neg_dtg * ((0.0 - dFlo_dYm) / dzc);
        di[i] = neg_dtg * ((dFhi_dYi - dFlo_dYi) / dzc) - 1.0;
        up[i] = neg_dtg * ((dFhi_dYp - 0.0) / dzc);
    }
}

void orc_compute_jacobian(const orc_problem *P, const orc_state *Y, const orc_cache *p,
                          double dtgamma, orc_jacobian *W)
{
    const int N = P->N;
    const int eh = (P->model == ORC_ENERGY_HYDROLOGY);
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        double *coef = (double *)malloc(sizeof(double) * (size_t)N * 2);
        double *A = coef + N;
        const double *K = (eh ? P->K_lag : p->K) + o;
        for (int i = 0; i < N; ++i) {
            double nu_eff = eh ? P->nu[o + i] - Y->theta_i[o + i] : P->nu[o + i];
            coef[i] = hcm_dpsidtheta(P, o + i, Y->theta_l[o + i], nu_eff);
        }
        double top_dflux = 0.0;
        /* haskey(p.soil, :dfluxBCdY): only a Richards cache with MoistureStateBC top has it */
        if (!eh && P->top_bc == ORC_TOP_MOISTURE_STATE) top_dflux = p->dfluxBCdY[c];
        column_tridiag_block(P, K, coef, dtgamma, top_dflux, W->w11_lo + o, W->w11_di + o, W->w11_up + o);
        if (eh) {
            /* (rho_e_int, theta_l): A = e_liq(T)*K, same coef, and "- I" (sic, energy_hydrology.jl:554-556) */
            for (int i = 0; i < N; ++i)
                A[i] = orc_volumetric_internal_energy_liq(p->T[o + i], P->rho_l, P->cp_l, P->T_ref) * K[i];
            column_tridiag_block(P, A, coef, dtgamma, 0.0, W->w21_lo + o, W->w21_di + o, W->w21_up + o);
            /* (rho_e_int, rho_e_int): A = kappa, coef = 1/rho_c_s(lagged theta_l, theta_i) */
            for (int i = 0; i < N; ++i)
                coef[i] = 1 / orc_volumetric_heat_capacity(P->theta_l_lag[o + i], Y->theta_i[o + i],
                                                           P->rho_c_ds[o + i], P->rho_l, P->cp_l,
                                                           P->rho_i, P->cp_i);
            column_tridiag_block(P, P->kappa_lag + o, coef, dtgamma, 0.0, W->w22_lo + o, W->w22_di + o,
                                 W->w22_up + o);
        }
        free(coef);
    }
}

/* ------------------------------------------------------------------ */
/* linear solve                                                        */
/* ------------------------------------------------------------------ */

/* Thomas sweep in the normalised (c', d') form ClimaCore's
 * MatrixFields single-field tridiagonal solve uses (un-vendored, ClimaCore 0.15.2;
 * restated from the published algorithm).  No pivoting: W is
 * strictly diagonally dominant for dtgamma >= 0. */
static void thomas(int N, const double *lo, const double *di, const double *up, const double *b,
                   double *x, double *cp)
{
    double den = 1.0 / di[0];
    cp[0] = up[0] * den;
    x[0] = b[0] * den;
    for (int i = 1; i < N; ++i) {
        den = 1.0 / (di[i] - lo[i] * cp[i - 1]);
        cp[i] = up[i] * den;
        x[i] = (b[i] - lo[i] * x[i - 1]) * den;
    }
    for (int i = N - 2; i >= 0; --i) x[i] = x[i] - cp[i] * x[i + 1];
}

/* src/shared_utilities/implicit_timestepping.jl:160-171:
 *  Richards: BlockDiagonalSolve -> Thomas on (theta_l,theta_l); x = -b for the -I blocks.
 *  EnergyHydrology: BlockLowerTriangularSolve(theta_l): solve W11 x1 = b1, then
 *  b2' = b2 - W21 x1, then W22 x2 = b2'; x = -b for theta_i and the two flux integrals.
 *  Call shape: test/integrated/full_land.jl:631-636. */
void orc_ldiv(const orc_problem *P, const orc_jacobian *W, const orc_state *b, orc_state *x)
{
    const int N = P->N;
    const int eh = (P->model == ORC_ENERGY_HYDROLOGY);
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        double *cp = (double *)malloc(sizeof(double) * (size_t)N * 2);
        double *b2 = cp + N;
        thomas(N, W->w11_lo + o, W->w11_di + o, W->w11_up + o, b->theta_l + o, x->theta_l + o, cp);
        x->intF_w[c] = -b->intF_w[c];
        if (eh) {
            const double *x1 = x->theta_l + o;
            for (int i = 0; i < N; ++i) {
                double s = W->w21_di[o + i] * x1[i];
                if (i > 0) s = W->w21_lo[o + i] * x1[i - 1] + s;
                if (i < N - 1) s = s + W->w21_up[o + i] * x1[i + 1];
                b2[i] = b->rho_e_int[o + i] - s;
            }
            thomas(N, W->w22_lo + o, W->w22_di + o, W->w22_up + o, b2, x->rho_e_int + o, cp);
            for (int i = 0; i < N; ++i) x->theta_i[o + i] = -b->theta_i[o + i];
            x->intF_e[c] = -b->intF_e[c];
        }
        free(cp);
    }
}

/* ------------------------------------------------------------------ */
/* Newton / ARS111 implicit stage                                      */
/* ------------------------------------------------------------------ */

static double *dalloc(int64_t n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

static void state_alloc(const orc_problem *P, orc_state *s)
{
    int64_t n3 = P->ncol * P->N;
    s->theta_l = dalloc(n3);
    s->rho_e_int = dalloc(n3);
    s->theta_i = dalloc(n3);
    s->intF_w = dalloc(P->ncol);
    s->intF_e = dalloc(P->ncol);
}

static void state_free(orc_state *s)
{
    free(s->theta_l);
    free(s->rho_e_int);
    free(s->theta_i);
    free(s->intF_w);
    free(s->intF_e);
}

/* ClimaTimeSteppers 0.10.6 (un-vendored) IMEX ARK stage with NewtonsMethod, as
 * configured at src/simulations/Simulations.jl:127-135 (ARS111, max_iters,
 * update_j = UpdateEvery(NewNewtonIteration), no convergence checker by default):
 *
 *   temp = U ; cache_imp!(U)
 *   for n in 1:max_iters
 *       Wfact(W, U, dtgamma)                    # compute_jacobian!
 *       f = T_imp!(U) ; f = temp + dtgamma*f - U
 *       dx = W \ f ; U -= dx
 *       converged && break
 *       n < max_iters && cache_imp!(U)
 *   end
 *
 * Restated from the published algorithm; parity unpinned against
 * the real package. */
int orc_implicit_step(const orc_problem *P, orc_state *U, orc_cache *p, orc_jacobian *W,
                      double dtgamma, int max_iters, double tol, double *last_dx_norm)
{
    const int N = P->N;
    const int eh = (P->model == ORC_ENERGY_HYDROLOGY);
    const int64_t n3 = P->ncol * N;
    orc_state temp, f, dx;
    state_alloc(P, &temp);
    state_alloc(P, &f);
    state_alloc(P, &dx);
    memcpy(temp.theta_l, U->theta_l, sizeof(double) * (size_t)n3);
    memcpy(temp.intF_w, U->intF_w, sizeof(double) * (size_t)P->ncol);
    if (eh) {
        memcpy(temp.rho_e_int, U->rho_e_int, sizeof(double) * (size_t)n3);
        memcpy(temp.theta_i, U->theta_i, sizeof(double) * (size_t)n3);
        memcpy(temp.intF_e, U->intF_e, sizeof(double) * (size_t)P->ncol);
    }
    orc_update_implicit_cache(P, U, p);
    int n = 0;
    double nrm = 0.0;
    for (n = 1; n <= max_iters; ++n) {
        orc_compute_jacobian(P, U, p, dtgamma, W);
        orc_compute_imp_tendency(P, U, p, &f);
        for (int64_t k = 0; k < n3; ++k) {
            f.theta_l[k] = temp.theta_l[k] + dtgamma * f.theta_l[k] - U->theta_l[k];
            if (eh) {
                f.rho_e_int[k] = temp.rho_e_int[k] + dtgamma * f.rho_e_int[k] - U->rho_e_int[k];
                f.theta_i[k] = temp.theta_i[k] + dtgamma * f.theta_i[k] - U->theta_i[k];
            }
        }
        for (int64_t c = 0; c < P->ncol; ++c) {
            f.intF_w[c] = temp.intF_w[c] + dtgamma * f.intF_w[c] - U->intF_w[c];
            if (eh) f.intF_e[c] = temp.intF_e[c] + dtgamma * f.intF_e[c] - U->intF_e[c];
        }
        orc_ldiv(P, W, &f, &dx);
        double ss = 0.0;
        for (int64_t k = 0; k < n3; ++k) {
            U->theta_l[k] -= dx.theta_l[k];
            ss += dx.theta_l[k] * dx.theta_l[k];
            if (eh) {
                U->rho_e_int[k] -= dx.rho_e_int[k];
                U->theta_i[k] -= dx.theta_i[k];
                ss += dx.rho_e_int[k] * dx.rho_e_int[k] + dx.theta_i[k] * dx.theta_i[k];
            }
        }
        for (int64_t c = 0; c < P->ncol; ++c) {
            U->intF_w[c] -= dx.intF_w[c];
            ss += dx.intF_w[c] * dx.intF_w[c];
            if (eh) {
                U->intF_e[c] -= dx.intF_e[c];
                ss += dx.intF_e[c] * dx.intF_e[c];
            }
        }
        nrm = sqrt(ss);
        if (tol >= 0.0 && nrm <= tol) break;
        if (n < max_iters) orc_update_implicit_cache(P, U, p);
    }
    if (n > max_iters) n = max_iters;
    if (last_dx_norm) *last_dx_norm = nrm;
    state_free(&temp);
    state_free(&f);
    state_free(&dx);
    return n;
}

/* ClimaCore column_integral_definite! of a centre field (rre.jl:502-511) */
void orc_column_integral(const orc_problem *P, const double *field, double *out)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += field[c * N + i] * dz_cell(P, i);
        out[c] = s;
    }
}

/* ------------------------------------------------------------------ */
/* explicit stage of EnergyHydrology: update_aux! and PhaseChange      */
/* (the step immediately before the implicit solve)  */
/* ------------------------------------------------------------------ */

/* soil_heat_parameterizations.jl:240-254 */
double orc_kappa_sat(double theta_l, double theta_i, double kappa_sat_unfrozen, double kappa_sat_frozen)
{
    double theta_w = theta_l + theta_i;
    if (theta_w < EPS64) return (kappa_sat_unfrozen + kappa_sat_frozen) / 2.0;
    return pow(kappa_sat_unfrozen, theta_l / theta_w) * pow(kappa_sat_frozen, theta_i / theta_w);
}

/* soil_heat_parameterizations.jl:283-285 */
double orc_relative_saturation(double theta_l, double theta_i, double nu) { return (theta_l + theta_i) / nu; }

/* soil_heat_parameterizations.jl:301-323 (Balland and Arp) */
double orc_kersten_number(double theta_i, double S_r, double alpha, double beta, double nu_ss_om,
                          double nu_ss_quartz, double nu_ss_gravel)
{
    if (theta_i < EPS64) {
        return pow(S_r, (1.0 + nu_ss_om - alpha * nu_ss_quartz - nu_ss_gravel) / 2.0) *
               pow(pow(1.0 + exp(-beta * S_r), -3.0) - pow((1.0 - S_r) / 2.0, 3.0), 1.0 - nu_ss_om);
    }
    return pow(S_r, 1.0 + nu_ss_om);
}

/* soil_heat_parameterizations.jl:266-269 */
double orc_thermal_conductivity(double kappa_dry, double K_e, double kappa_sat)
{
    return K_e * kappa_sat + (1.0 - K_e) * kappa_dry;
}

/* soil_heat_parameterizations.jl:59-61: 3 * rho_c * dz^2 / kappa */
double orc_thermal_time(double rho_c, double dz, double kappa) { return 3.0 * rho_c * (dz * dz) / kappa; }

static inline double any_matric_potential(int closure, double a, double b, double m, double S)
{
    return closure == ORC_VAN_GENUCHTEN ? orc_vg_matric_potential(a, b, m, S) : orc_bc_matric_potential(a, b, S);
}

static inline double any_inverse_matric_potential(int closure, double a, double b, double m, double psi)
{
    return closure == ORC_VAN_GENUCHTEN ? orc_vg_inverse_matric_potential(a, b, m, psi)
                                        : orc_bc_inverse_matric_potential(a, b, psi);
}

/* soil_heat_parameterizations.jl:35-52.  The signature is (.., _rho_ice, _rho_liq, ..) and the body
 * uses _rho_ice / _rho_liq; the two arguments are named by POSITION here because the reference's two call
 * sites pass them in different orders (see orc_update_aux). */
double orc_soil_Tf_depressed(int closure, double a, double b, double m, double theta_l, double theta_i, double nu,
                             double theta_r, double rho_first, double rho_second, double T_freeze, double grav,
                             double LH_f0)
{
    double theta_tot = dmin(rho_first / rho_second * theta_i + theta_l, nu);
    double psi_w0 = any_matric_potential(closure, a, b, m, orc_effective_saturation(nu, theta_tot, theta_r));
    return dmax(T_freeze * exp(grav * psi_w0 / LH_f0), 1.0);
}

/* soil_heat_parameterizations.jl:89-122 */
double orc_phase_change_source(int closure, double a, double b, double m, double theta_l, double theta_i, double T,
                               double tau, double nu, double theta_r, double rho_i, double rho_l, double LH_f0,
                               double T_freeze, double grav)
{
    double Tf = orc_soil_Tf_depressed(closure, a, b, m, theta_l, theta_i, nu, theta_r, rho_i, rho_l, T_freeze, grav,
                                      LH_f0);
    double psi_T = LH_f0 / grav * log(T / Tf) * orc_heaviside(Tf - T, 0.0);
    double theta_tot = dmin(rho_i / rho_l * theta_i + theta_l, nu);
    double psi_w0 = any_matric_potential(closure, a, b, m, orc_effective_saturation(nu, theta_tot, theta_r));
    double theta_star = any_inverse_matric_potential(closure, a, b, m, psi_w0 + psi_T) * (nu - theta_r) + theta_r;
    return (theta_l - theta_star) / tau;
}

/* update_aux!(p, Y, t): energy_hydrology.jl:722-814.
 *   theta_l  :746-747   volumetric_liquid_fraction(theta_l, nu - theta_i, theta_r)
 *   kappa    :757-771   thermal_conductivity(kappa_dry, kersten_number(..), kappa_sat(..))
 *   T        :773-783   temperature_from_rho_e_int(rho_e_int, theta_i, volumetric_heat_capacity(p.theta_l, ..))
 *   K        :785-792   impedance * viscosity * hydraulic_conductivity(effective_saturation(nu, theta_l, theta_r))
 *                       -- the saturation uses nu, NOT nu - theta_i, and the augmented theta_l (Y), not p.theta_l
 *   psi      :793-794   pressure_head(.., nu - theta_i, ..)
 *   Tf_depressed :800-811  called as soil_Tf_depressed(.., _rho_l, _rho_i, ..) while the function's positional
 *                       parameters are (_rho_ice, _rho_liq): the cached field therefore uses rho_l / rho_i as
 *                       the ice-to-liquid ratio.  PhaseChange (below) calls it with (_rho_i, _rho_l).  Both are
 *                       restated as the reference evaluates them.
 *   total_water  :1292-1306  column integral of theta_l + theta_i * rho_i / rho_l
 *   total_energy :1318-1327  column integral of rho_e_int */
void orc_update_aux(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, orc_aux *a)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        double tw = 0.0, te = 0.0;
        for (int i = 0; i < N; ++i) {
            const int64_t k = c * N + i;
            const double nu = P->nu[k], theta_r = P->theta_r[k];
            const double th = Y->theta_l[k], thi = Y->theta_i[k];
            const double theta_l = orc_volumetric_liquid_fraction(th, nu - thi, theta_r);
            a->theta_l[k] = theta_l;
            a->kappa[k] = orc_thermal_conductivity(
                X->kappa_dry[k],
                orc_kersten_number(thi, orc_relative_saturation(theta_l, thi, nu), X->alpha, X->beta, X->nu_ss_om[k],
                                   X->nu_ss_quartz[k], X->nu_ss_gravel[k]),
                orc_kappa_sat(theta_l, thi, X->kappa_sat_unfrozen[k], X->kappa_sat_frozen[k]));
            const double T = orc_temperature_from_rho_e_int(
                Y->rho_e_int[k], thi,
                orc_volumetric_heat_capacity(theta_l, thi, P->rho_c_ds[k], P->rho_l, P->cp_l, P->rho_i, P->cp_i),
                P->rho_i, P->T_ref, P->LH_f0);
            a->T[k] = T;
            a->K[k] = orc_impedance_factor(thi / (theta_l + thi - theta_r), X->Omega) *
                      orc_viscosity_factor(T, X->gamma, X->gammaT_ref) *
                      hcm_conductivity(P, k, orc_effective_saturation(nu, th, theta_r));
            a->psi[k] = hcm_pressure_head(P, k, th, nu - thi);
            a->Tf_depressed[k] = orc_soil_Tf_depressed(P->closure, P->hcm_a[k], P->hcm_b[k], P->hcm_m ? P->hcm_m[k] : 0.0,
                                                       theta_l, thi, nu, theta_r, P->rho_l, P->rho_i, X->T_freeze,
                                                       X->grav, P->LH_f0);
            tw += (th + thi * P->rho_i / P->rho_l) * dz_cell(P, i);
            te += Y->rho_e_int[k] * dz_cell(P, i);
        }
        if (a->total_water) a->total_water[c] = tw;
        if (a->total_energy) a->total_energy[c] = te;
    }
}

/* source!(dY, ::PhaseChange, Y, p, model): energy_hydrology.jl:846-906; dz = layer thickness
 * (model.domain.fields.dz = Domains.get_dz's third value, Domains.jl:947-948) */
void orc_phase_change(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, const orc_aux *a,
                      double *dtheta_l, double *dtheta_i)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        for (int i = 0; i < N; ++i) {
            const int64_t k = c * N + i;
            const double thi = Y->theta_i[k];
            const double rho_c = orc_volumetric_heat_capacity(a->theta_l[k], thi, P->rho_c_ds[k], P->rho_l, P->cp_l,
                                                              P->rho_i, P->cp_i);
            const double tau = orc_thermal_time(rho_c, dz_cell(P, i), a->kappa[k]);
            const double s = orc_phase_change_source(P->closure, P->hcm_a[k], P->hcm_b[k], P->hcm_m ? P->hcm_m[k] : 0.0,
                                                     a->theta_l[k], thi, a->T[k], tau, P->nu[k], P->theta_r[k],
                                                     P->rho_i, P->rho_l, P->LH_f0, X->T_freeze, X->grav);
            dtheta_l[k] += -s;
            dtheta_i[k] += (P->rho_l / P->rho_i) * s;
        }
    }
}

/* ------------------------------------------------------------------ */
/* TOPMODEL runoff (explicit stage): produces the lagged inputs of the */
/* implicit TOPMODELSubsurfaceRunoff source (R_ss, R_ess, h_grad,      */
/* is_saturated) and the infiltration / surface runoff                 */
/* ------------------------------------------------------------------ */

/* Runoff.jl:421-423 */
double orc_topmodel_ss_flux(double R_sb, double f_over, double z_wt) { return R_sb * exp(-f_over * z_wt); }

/* Runoff.jl:373-376 */
double orc_topmodel_surface_infiltration(double f_max, double f_over, double z_wt, double f_ic, double precip)
{
    double f_sat = dmin(f_max * exp(-f_over / 2.0 * z_wt), 1.0);
    return (1.0 - f_sat) * dmax(f_ic, precip);
}

/* Runoff.jl:234-283 (+ soil_infiltration_capacity :385-410, is_saturated :432-434,
 * update_subsurface_energy_runoff! :266-279) */
void orc_update_runoff(const orc_problem *P, const orc_explicit_params *X, const orc_runoff_params *R, const orc_state *Y,
                       const orc_aux *a, const double *precip, orc_runoff *out)
{
    const int N = P->N;
    const int eh = P->model == ORC_ENERGY_HYDROLOGY;
    FOR_COLUMNS(P, c)
    {
        double h_all = 0.0, h_liq = 0.0, e_liq = 0.0;
        for (int i = 0; i < N; ++i) {
            const int64_t k = c * N + i;
            const double th = Y->theta_l[k], thi = eh ? Y->theta_i[k] : 0.0;
            const double nu = P->nu[k], theta_r = P->theta_r[k];
            /* weighted by how much above saturation, ice included (:251-253) */
            const double s_all = orc_heaviside((th + thi - theta_r) - (nu - theta_r), 0.0) * (th + thi - theta_r) / (nu - theta_r);
            /* liquid water only, for the subsurface runoff (:259-261) */
            const double s_liq = orc_heaviside((th - theta_r) - (nu - theta_r), 0.0) * (th - theta_r) / (nu - theta_r);
            out->is_saturated[k] = s_liq;
            h_all += s_all * dz_cell(P, i);
            h_liq += s_liq * dz_cell(P, i);
            if (eh) e_liq += s_liq * orc_volumetric_internal_energy_liq(a->T[k], P->rho_l, P->cp_l, P->T_ref) * dz_cell(P, i);
        }
        /* infiltration capacity at the top centre (:385-410) */
        const int64_t kt = c * N + N - 1;
        double ic = -1 * P->K_sat[kt];
        if (eh)
            ic = -P->K_sat[kt] *
                 orc_impedance_factor(Y->theta_i[kt] / (a->theta_l[kt] + Y->theta_i[kt] - P->theta_r[kt]), X->Omega) *
                 orc_viscosity_factor(a->T[kt], X->gamma, X->gammaT_ref);
        out->infiltration[c] = orc_topmodel_surface_infiltration(R->f_max[c], R->f_over, R->depth - h_all, ic, precip[c]);
        out->R_s[c] = fabs(precip[c] - out->infiltration[c]);
        out->h_grad[c] = h_liq;
        out->R_ss[c] = orc_topmodel_ss_flux(R->R_sb, R->f_over, R->depth - h_liq);
        if (eh && out->R_ess) out->R_ess[c] = e_liq * (out->R_ss[c] / dmax(h_liq, EPS64));
    }
}

/* ------------------------------------------------------------------ */
/* Atmosphere-driven top boundary fluxes (second half) */
/* ------------------------------------------------------------------ */

/* soil_infiltration_capacity: Runoff.jl:385-410 (RichardsModel: -K_sat; EnergyHydrology: -K_sat x impedance x
 * viscosity at the top cell centre, from p.soil.theta_l / T) */
static double infiltration_capacity(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y,
                                    const orc_aux *a, int64_t kt)
{
    if (P->model != ORC_ENERGY_HYDROLOGY) return -1 * P->K_sat[kt];
    return -P->K_sat[kt] *
           orc_impedance_factor(Y->theta_i[kt] / (a->theta_l[kt] + Y->theta_i[kt] - P->theta_r[kt]), X->Omega) *
           orc_viscosity_factor(a->T[kt], X->gamma, X->gammaT_ref);
}

/* update_infiltration_water_flux!(p, ::NoRunoff, input, ...): Runoff.jl:69-71 (infiltration = input) and
 * update_infiltration_water_flux!(p, ::SurfaceRunoff, input, Y, t, model): Runoff.jl:129-148 with
 * surface_infiltration :109-111 and is_saturated :432-434 (heaviside(theta_l + theta_i - nu) per cell).
 * kind: 0 NoRunoff, 1 SurfaceRunoff.  is_saturated / R_s may be NULL for NoRunoff. */
void orc_surface_runoff(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, const orc_aux *a,
                        int kind, const double *input, double *is_saturated, double *infiltration, double *R_s)
{
    const int N = P->N;
    const int eh = P->model == ORC_ENERGY_HYDROLOGY;
    FOR_COLUMNS(P, c)
    {
        if (kind == 0) {
            infiltration[c] = input[c];
            continue;
        }
        for (int i = 0; i < N; ++i) {
            const int64_t k = c * N + i;
            is_saturated[k] = orc_heaviside((Y->theta_l[k] + (eh ? Y->theta_i[k] : 0.0)) - P->nu[k], 0.0);
        }
        const int64_t kt = c * N + N - 1;
        const double ic = infiltration_capacity(P, X, Y, a, kt);
        infiltration[c] = (1 - is_saturated[kt]) * dmax(ic, input[c]);
        R_s[c] = fabs(input[c] - infiltration[c]);
    }
}

/* soil_boundary_fluxes!(bc::AtmosDrivenFluxBC, Val((:soil,)), model, Y, p, t): boundary_conditions.jl:901-936 after the
 * runoff has partitioned the liquid influx (compute_liquid_influx :955-961 = p.drivers.P_liq):
 *   top_bc.water = infiltration + turbulent_fluxes.vapor_flux_liq
 *   top_bc.heat  = R_n + lhf + shf + infiltration * volumetric_internal_energy_liq(p.drivers.T)   (:988-1002)
 * The turbulent fluxes and the net radiation are the host model's (SurfaceFluxes.jl, radiation drivers). */
void orc_atmos_driven_top_fluxes(const orc_problem *P, const double *infiltration, const double *vapor_flux_liq,
                                 const double *lhf, const double *shf, const double *R_n, const double *T_air,
                                 double *top_bc_w, double *top_bc_h)
{
    FOR_COLUMNS(P, c)
    {
        top_bc_w[c] = infiltration[c] + vapor_flux_liq[c];
        const double inf_e = infiltration[c] * orc_volumetric_internal_energy_liq(T_air[c], P->rho_l, P->cp_l, P->T_ref);
        top_bc_h[c] = R_n[c] + lhf[c] + shf[c] + inf_e;
    }
}

/* soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...): boundary_conditions.jl:590-608:
 *   bottom_bc.water = -K_1, bottom_bc.heat = -K_1 * volumetric_internal_energy_liq(T_1)  (level 1 = the bottom cell) */
void orc_energy_water_free_drainage(const orc_problem *P, const orc_aux *a, double *bot_bc_w, double *bot_bc_h)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        const int64_t k = c * N;
        bot_bc_w[c] = -1 * a->K[k];
        bot_bc_h[c] = -1 * a->K[k] * orc_volumetric_internal_energy_liq(a->T[k], P->rho_l, P->cp_l, P->T_ref);
    }
}

/* ------------------------------------------------------------------ */
/* SoilCO2Model: implicit CO2 / O2 diffusion        */
/* ------------------------------------------------------------------ */

/* boundary_flux!(..., ::AtmosCO2StateBC / ::AtmosO2StateBC, ::TopBoundary, ...): Biogeochemistry.jl:932-957,
 * 1078-1111: diffusive_flux(D_N, c_atm, max(C_N / theta_N, 0), dz_top) and dfluxBCdY = D_N / theta_N / dz_top */
void orc_co2_boundary_flux(const orc_problem *P, const orc_co2_species *S, const double *C, double *top_bc,
                           double *dfluxBCdY)
{
    const int N = P->N;
    if (!S->c_atm) return;
    FOR_COLUMNS(P, c)
    {
        const int64_t k = c * N + N - 1;
        const double dz = dz_top(P);
        top_bc[c] = -S->D[k] * (S->c_atm[c] - dmax(C[k] / S->theta_eff[k], 0.0)) / dz;
        if (dfluxBCdY) dfluxBCdY[c] = S->D[k] / S->theta_eff[k] / dz;
    }
}

/* compute_imp_tendency!: Biogeochemistry.jl:371-413:
 *   dC = -div(-interp(D) grad(max(C, 0) / theta_eff)), boundary faces = top_bc / bottom_bc */
void orc_co2_imp_tendency(const orc_problem *P, const orc_co2_species *S, const double *C, const double *top_bc,
                          const double *bot_bc, double *dC)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        double q_lo = bot_bc[c];
        for (int i = 0; i < N; ++i) {
            double q_hi;
            if (i < N - 1) {
                const double u0 = dmax(C[o + i], 0.0) / S->theta_eff[o + i];
                const double u1 = dmax(C[o + i + 1], 0.0) / S->theta_eff[o + i + 1];
                q_hi = -((S->D[o + i] + S->D[o + i + 1]) / 2.0) * ((u1 - u0) / dz_face(P, i + 1));
            } else {
                q_hi = top_bc[c];
            }
            dC[o + i] = -((q_hi - q_lo) / dz_cell(P, i));
            q_lo = q_hi;
        }
    }
}

/* compute_jacobian!: Biogeochemistry.jl:1119-1195: dtgamma (D . (Diag(interp(D)) . G . Diag(1/theta_eff)
 * - Lower(dfluxBCdY))) - I: the block of column_tridiag_block with A = D, coef = 1/theta_eff */
void orc_co2_jacobian(const orc_problem *P, const orc_co2_species *S, double dtgamma, const double *dfluxBCdY,
                      double *lo, double *di, double *up)
{
    const int N = P->N;
    FOR_COLUMNS(P, c)
    {
        const int64_t o = c * N;
        double *coef = (double *)malloc(sizeof(double) * (size_t)N);
        for (int i = 0; i < N; ++i) coef[i] = 1 / S->theta_eff[o + i];
        column_tridiag_block(P, S->D + o, coef, dtgamma, (S->c_atm && dfluxBCdY) ? dfluxBCdY[c] : 0.0, lo + o, di + o,
                             up + o);
        free(coef);
    }
}

int orc_co2_implicit_step(const orc_problem *P, const orc_co2_species *S, double *C, double *top_bc,
                          const double *bot_bc, double dtgamma, int max_iters)
{
    const int N = P->N;
    const int64_t n3 = P->ncol * (int64_t)N;
    double *temp = (double *)malloc(sizeof(double) * (size_t)n3 * 6);
    double *f = temp + n3, *lo = f + n3, *di = lo + n3, *up = di + n3, *x = up + n3;
    double *dflux = (double *)calloc((size_t)P->ncol, sizeof(double));
    memcpy(temp, C, sizeof(double) * (size_t)n3);
    for (int it = 0; it < max_iters; ++it) {
        orc_co2_boundary_flux(P, S, C, top_bc, dflux);       /* update_implicit_boundary_fluxes! (:320-357) */
        orc_co2_jacobian(P, S, dtgamma, dflux, lo, di, up);
        orc_co2_imp_tendency(P, S, C, top_bc, bot_bc, f);
        for (int64_t k = 0; k < n3; ++k) f[k] = temp[k] + dtgamma * f[k] - C[k];
        FOR_COLUMNS(P, c)
        {
            double *cp = (double *)malloc(sizeof(double) * (size_t)N);
            thomas(N, lo + c * N, di + c * N, up + c * N, f + c * N, x + c * N, cp);
            free(cp);
        }
        for (int64_t k = 0; k < n3; ++k) C[k] -= x[k];
    }
    free(temp);
    free(dflux);
    return max_iters;
}
