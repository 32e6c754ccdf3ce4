/*
 * soil_oracle.h -- CPU restatement (plain C, FP64) of ClimaLand.jl's implicit
 * soil-column path.  TEST INFRASTRUCTURE ONLY: nothing under the product
 * package may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the timed CPU baseline, never as the shipped path.
 *
 * PARITY STATUS: the point functions, the implicit tendency stencil and the
 * Jacobian entries are pinned against the known-answer tests the reference
 * holds (tests/test_oracle_reference_kats.py lists each reference test
 * file:line).  The LINEAR SOLVE and the NEWTON/ARS111 STEP live in un-vendored
 * dependencies (ClimaCore 0.15.2 MatrixFields.field_matrix_solve!,
 * ClimaTimeSteppers 0.10.6 NewtonsMethod; pinned in
 * /root/reference/.buildkite/Manifest.toml:510-514,562-566) and no Julia
 * toolchain exists in the build image, so for those two rows the oracle is a
 * restatement of the published algorithm: "parity unpinned" against the real
 * reference for solve/step (checked instead by residual, scipy solve_banded,
 * mass balance).
 *
 * Layout: the reference's own CPU layout, level fastest: a[c*N + i],
 * i = 0..N-1 bottom -> top (Fields.level(.,1) is the bottom,
 * src/standalone/Soil/boundary_conditions.jl:351).  z <= 0, increasing upward.
 * Fluxes positive in +z.
 */
#ifndef SOIL_ORACLE_H
#define SOIL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RICHARDS = 0, ORC_ENERGY_HYDROLOGY = 1 };
enum { ORC_VAN_GENUCHTEN = 0, ORC_BROOKS_COREY = 1 };
enum { ORC_TOP_FLUX = 0, ORC_TOP_MOISTURE_STATE = 1 };
enum { ORC_BOT_FLUX = 0, ORC_BOT_FREE_DRAINAGE = 1, ORC_BOT_MOISTURE_STATE = 2 };

/* Everything the implicit path reads.  Per-cell arrays are [ncol*N]
 * (level fastest); per-column arrays are [ncol]. */
typedef struct {
    int32_t model, closure, top_bc, bottom_bc, has_topmodel_source;
    int32_t N;
    int64_t ncol;
    int32_t nthreads;          /* OpenMP threads over columns (<=1: serial) */
    const double *z_c;         /* [N]   cell centres */
    const double *z_f;         /* [N+1] cell faces   */
    /* hydrology parameters (RichardsParameters / EnergyHydrologyParameters) */
    const double *nu, *theta_r, *K_sat, *S_s;
    const double *hcm_a;       /* vG alpha | BC c    */
    const double *hcm_b;       /* vG n     | BC psi_b */
    const double *hcm_m;       /* vG m     | unused  */
    /* EnergyHydrology only */
    const double *rho_c_ds;
    const double *K_lag, *kappa_lag, *theta_l_lag;   /* lagged p.soil.{K,kappa,theta_l} */
    /* implicit TOPMODEL source (lagged) */
    const double *is_saturated;                      /* [ncol*N] */
    const double *R_ss, *R_ess, *h_grad;             /* [ncol] */
    /* state boundary values theta_bc (MoistureStateBC), [ncol] */
    const double *theta_bc_top, *theta_bc_bot;
    /* LandParameters constants (src/shared_utilities/Parameters.jl:11-58) */
    double rho_l, rho_i, cp_l, cp_i, T_ref, LH_f0;
} orc_problem;

/* prognostic state Y (rre.jl:282, energy_hydrology.jl prognostic_vars) */
typedef struct {
    double *theta_l;      /* [ncol*N] */
    double *rho_e_int;    /* [ncol*N] EH only */
    double *theta_i;      /* [ncol*N] EH only */
    double *intF_w;       /* [ncol]  integral of boundary water flux */
    double *intF_e;       /* [ncol]  EH only */
} orc_state;

/* the part of the cache p.soil the implicit path writes / reads */
typedef struct {
    double *K, *psi, *T;              /* [ncol*N] */
    double *top_bc_w, *bot_bc_w;      /* [ncol] boundary water fluxes */
    double *top_bc_h, *bot_bc_h;      /* [ncol] boundary heat fluxes (EH) */
    double *dfluxBCdY;                /* [ncol] only for ORC_TOP_MOISTURE_STATE */
    double *total_water;              /* [ncol] Richards diagnostic */
} orc_cache;

/* Jacobian blocks, TridiagonalMatrixRow per cell: lower/diag/upper */
typedef struct {
    double *w11_lo, *w11_di, *w11_up;   /* (theta_l, theta_l) */
    double *w21_lo, *w21_di, *w21_up;   /* (rho_e_int, theta_l)  EH */
    double *w22_lo, *w22_di, *w22_up;   /* (rho_e_int, rho_e_int) EH */
} orc_jacobian;

/* ---- point functions (soil_hydrology_parameterizations.jl, soil_heat_parameterizations.jl) */
double orc_effective_saturation(double nu_eff, double theta_l, double theta_r);
double orc_volumetric_liquid_fraction(double theta_l, double nu_eff, double theta_r);
double orc_vg_matric_potential(double alpha, double n, double m, double S);
double orc_vg_inverse_matric_potential(double alpha, double n, double m, double psi);
double orc_vg_pressure_head(double alpha, double n, double m, double theta_r, double theta_l, double nu_eff, double S_s);
double orc_vg_dpsidtheta(double alpha, double n, double m, double theta, double nu_eff, double theta_r, double S_s);
double orc_vg_hydraulic_conductivity(double m, double K_sat, double S);
double orc_bc_matric_potential(double c, double psi_b, double S);
double orc_bc_inverse_matric_potential(double c, double psi_b, double psi);
double orc_bc_pressure_head(double c, double psi_b, double theta_r, double theta_l, double nu_eff, double S_s);
double orc_bc_dpsidtheta(double c, double psi_b, double theta, double nu_eff, double theta_r, double S_s);
double orc_bc_hydraulic_conductivity(double c, double K_sat, double S);
double orc_volumetric_heat_capacity(double theta_l, double theta_i, double rho_c_ds, double rho_l, double cp_l, double rho_i, double cp_i);
double orc_temperature_from_rho_e_int(double rho_e_int, double theta_i, double rho_c_s, double rho_i, double T_ref, double LH_f0);
double orc_volumetric_internal_energy(double theta_i, double rho_c_s, double T, double rho_i, double T_ref, double LH_f0);
double orc_volumetric_internal_energy_liq(double T, double rho_l, double cp_l, double T_ref);
double orc_heaviside(double x, double a);
double orc_impedance_factor(double f_i, double Omega);
double orc_viscosity_factor(double T, double gamma, double gammaT_ref);

double orc_kappa_sat(double theta_l, double theta_i, double kappa_sat_unfrozen, double kappa_sat_frozen);
double orc_relative_saturation(double theta_l, double theta_i, double nu);
double orc_kersten_number(double theta_i, double S_r, double alpha, double beta, double nu_ss_om,
                          double nu_ss_quartz, double nu_ss_gravel);
double orc_thermal_conductivity(double kappa_dry, double K_e, double kappa_sat);
double orc_thermal_time(double rho_c, double dz, double kappa);
/* closure: ORC_VAN_GENUCHTEN (a = alpha, b = n, m) | ORC_BROOKS_COREY (a = c, b = psi_b).  rho_first / rho_second
 * are the reference's positional `_rho_ice`, `_rho_liq` arguments (see orc_update_aux for why they are named
 * by position). */
double orc_soil_Tf_depressed(int closure, double a, double b, double m, double theta_l, double theta_i, double nu,
                             double theta_r, double rho_first, double rho_second, double T_freeze, double grav,
                             double LH_f0);
double orc_phase_change_source(int closure, double a, double b, double m, double theta_l, double theta_i, double T,
                               double tau, double nu, double theta_r, double rho_i, double rho_l, double LH_f0,
                               double T_freeze, double grav);

/* ---- explicit stage of EnergyHydrology ------------------------------------- */
/* the parameters only the explicit stage reads (EnergyHydrologyParameters, energy_hydrology.jl:60-170) */
typedef struct {
    const double *kappa_dry, *kappa_sat_unfrozen, *kappa_sat_frozen;   /* [ncol*N] */
    const double *nu_ss_om, *nu_ss_quartz, *nu_ss_gravel;              /* [ncol*N] */
    double Omega, gamma, gammaT_ref, alpha, beta;                      /* scalars (:150-160) */
    double T_freeze, grav;                                             /* LandParameters */
} orc_explicit_params;

/* p.soil.{theta_l, kappa, T, K, psi, Tf_depressed, total_water, total_energy} */
typedef struct {
    double *theta_l, *kappa, *T, *K, *psi, *Tf_depressed;   /* [ncol*N] */
    double *total_water, *total_energy;                     /* [ncol] */
} orc_aux;

/* update_aux!(p, Y, t) of EnergyHydrology: energy_hydrology.jl:722-814 */
void orc_update_aux(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, orc_aux *a);
/* source!(dY, ::PhaseChange, Y, p, model): energy_hydrology.jl:846-906; ADDS into dtheta_l / dtheta_i */
void orc_phase_change(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, const orc_aux *a,
                      double *dtheta_l, double *dtheta_i);

/* ---- TOPMODEL runoff of the explicit stage: Runoff/Runoff.jl:234-283, 373-451 */
double orc_topmodel_ss_flux(double R_sb, double f_over, double z_wt);
double orc_topmodel_surface_infiltration(double f_max, double f_over, double z_wt, double f_ic, double precip);
typedef struct {
    const double *f_max;      /* [ncol] */
    double f_over, R_sb, depth;
} orc_runoff_params;
typedef struct {
    double *is_saturated;                                   /* [ncol*N] (the liquid-water weighting, :259-260) */
    double *h_grad, *infiltration, *R_s, *R_ss, *R_ess;     /* [ncol] */
} orc_runoff;
/* update_infiltration_water_flux!(p, ::TOPMODELRunoff, input, Y, t, model): Runoff.jl:234-283.  X and a (Omega,
 * gamma, gammaT_ref; p.soil.theta_l, T) are read for EnergyHydrology only and may be NULL for RichardsModel. */
void orc_update_runoff(const orc_problem *P, const orc_explicit_params *X, const orc_runoff_params *R, const orc_state *Y,
                       const orc_aux *a, const double *precip, orc_runoff *out);

/* ---- atmosphere-driven top boundary fluxes: boundary_conditions.jl:590-608, 901-1002,
 *      Runoff/Runoff.jl:69-71, 109-148.  kind: 0 NoRunoff, 1 SurfaceRunoff (TOPMODELRunoff: orc_update_runoff). */
void orc_surface_runoff(const orc_problem *P, const orc_explicit_params *X, const orc_state *Y, const orc_aux *a,
                        int kind, const double *input, double *is_saturated, double *infiltration, double *R_s);
void orc_atmos_driven_top_fluxes(const orc_problem *P, const double *infiltration, const double *vapor_flux_liq,
                                 const double *lhf, const double *shf, const double *R_n, const double *T_air,
                                 double *top_bc_w, double *top_bc_h);
void orc_energy_water_free_drainage(const orc_problem *P, const orc_aux *a, double *bot_bc_w, double *bot_bc_h);

/* ---- SoilCO2Model implicit diffusion: Biogeochemistry.jl:320-413, 1119-1195 ---------- */
/* One diffusing species (CO2 with p.soilco2.{D, theta_eff}, or O2 with {D_o2, theta_eff_o2}); both have the same
 * G . Diag . D structure.  c_atm != NULL: the top BC is the atmosphere's state (AtmosCO2StateBC / AtmosO2StateBC,
 * :932-957, :1078-1111), re-evaluated every Newton iteration together with dfluxBCdY; else top_bc is a flux value. */
typedef struct {
    const double *D, *theta_eff;   /* [ncol*N] lagged */
    const double *c_atm;           /* [ncol] air-equivalent concentration at the surface, or NULL */
} orc_co2_species;
void orc_co2_boundary_flux(const orc_problem *P, const orc_co2_species *S, const double *C, double *top_bc,
                           double *dfluxBCdY);
void orc_co2_imp_tendency(const orc_problem *P, const orc_co2_species *S, const double *C, const double *top_bc,
                          const double *bot_bc, double *dC);
void orc_co2_jacobian(const orc_problem *P, const orc_co2_species *S, double dtgamma, const double *dfluxBCdY,
                      double *lo, double *di, double *up);
/* Newton loop of one implicit ARS111 stage on C (in: temp, out: new value); top_bc is updated in place for a
 * state BC.  Returns the iterations done. */
int orc_co2_implicit_step(const orc_problem *P, const orc_co2_species *S, double *C, double *top_bc,
                          const double *bot_bc, double dtgamma, int max_iters);

/* ---- the hooks (each cites the reference in soil_oracle.c) */
void orc_update_implicit_cache(const orc_problem *P, const orc_state *Y, orc_cache *p);
/* explicit-stage flavour of the boundary-flux update: always evaluates (rre.jl:111-149) */
void orc_update_boundary_fluxes(const orc_problem *P, const orc_state *Y, orc_cache *p);
void orc_compute_imp_tendency(const orc_problem *P, const orc_state *Y, const orc_cache *p, orc_state *dY);
void orc_compute_jacobian(const orc_problem *P, const orc_state *Y, const orc_cache *p, double dtgamma, orc_jacobian *W);
/* x = W^{-1} b   (BlockDiagonalSolve / BlockLowerTriangularSolve(theta_l); x = -b for -I blocks) */
void orc_ldiv(const orc_problem *P, const orc_jacobian *W, const orc_state *b, orc_state *x);
/* one implicit ARS111 stage: Newton loop of SURVEY 3.2 on U (in: temp, out: U_new).
 * Returns iterations done.  tol < 0: fixed max_iters (reference default).  tol >= 0:
 * stop when ||dx||_2 over all columns and fields <= tol.  last_dx_norm may be NULL. */
int orc_implicit_step(const orc_problem *P, orc_state *U, orc_cache *p, orc_jacobian *W,
                      double dtgamma, int max_iters, double tol, double *last_dx_norm);
/* sum_i theta_i * dz_c[i] per column (rre.jl:502-511) */
void orc_column_integral(const orc_problem *P, const double *field, double *out);

#ifdef __cplusplus
}
#endif
#endif
