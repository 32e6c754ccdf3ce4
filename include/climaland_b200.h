/*
 * climaland_b200.h -- C ABI of libclimaland_b200.so: the B200 (sm_100a, FP64)
 * implementation of ClimaLand.jl's implicit soil-column path.
 *
 * This is the drop-in boundary.  ClimaLand's Julia hooks keep their
 * names and call these entry points through `ccall` (binding shown in
 * INTEGRATION.md and julia/ClimaLandB200.jl).  Each entry point cites the
 * reference interface it replaces; paths are relative to the ClimaLand.jl tree.
 *
 *   make_update_implicit_cache  src/shared_utilities/models.jl:238-246
 *   make_compute_imp_tendency   src/standalone/Soil/rre.jl:161-203,
 *                               src/standalone/Soil/energy_hydrology.jl:363-425
 *   make_compute_jacobian       rre.jl:391-458, energy_hydrology.jl:466-576
 *   initialize_jacobian + ldiv! src/shared_utilities/implicit_timestepping.jl:63-172
 *                               (ClimaCore MatrixFields.field_matrix_solve!)
 *   Newton/ARS111 stage         src/simulations/Simulations.jl:127-135
 *                               (ClimaTimeSteppers IMEXAlgorithm + NewtonsMethod)
 *   make_update_aux (EH)        energy_hydrology.jl:722-814   } the explicit stage right before the
 *   source!(::PhaseChange)      energy_hydrology.jl:846-906   } implicit solve
 *
 * Conventions
 *   - plain C: pointers, sizes, strides; no C++ or torch types.
 *   - every function returns CLB_OK (0) or a negative clb_status; the message is
 *     available from clb_last_error().  Nothing throws or exits across the ABI.
 *   - levels i = 0..N-1 run bottom -> top (Fields.level(.,1) is the bottom,
 *     src/standalone/Soil/boundary_conditions.jl:351); z <= 0 increases upward;
 *     fluxes are positive in +z.
 *   - a handle is bound to one CUDA device and one stream; calls on a handle are
 *     serialised by the caller (ClimaTimeSteppers drives the hooks from one task).
 *     Work is enqueued asynchronously on the stream unless stated otherwise.
 *   - the library keeps its own column-fastest (SoA over levels) device mirrors:
 *     element (level i, column c) of a per-cell field lives at  base[i*ld + c].
 *     Callers never see that layout: clb_set_field / clb_get_field take the
 *     caller's own strides (ClimaCore `parent(field)` is level-fastest:
 *     stride_level = 1, stride_column = N*Nf) and transpose on the device.
 *   - there is no CPU fallback: without a CUDA device clb_create fails.
 */
#ifndef CLIMALAND_B200_H
#define CLIMALAND_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define CLB_ABI_VERSION 3

typedef struct clb_handle_s *clb_handle;

typedef enum {
    CLB_OK = 0,
    CLB_ERR_INVALID = -1,      /* bad argument / configuration            */
    CLB_ERR_CUDA = -2,         /* CUDA runtime error (see clb_last_error) */
    CLB_ERR_UNSET = -3,        /* a field the hook needs was never set    */
    CLB_ERR_NCCL = -4,         /* NCCL missing or failed                  */
    CLB_ERR_NO_DEVICE = -5     /* no CUDA device: there is no CPU path    */
} clb_status;

/* model kinds: RichardsModel (rre.jl) / EnergyHydrology (energy_hydrology.jl) */
enum { CLB_RICHARDS = 0, CLB_ENERGY_HYDROLOGY = 1 };
/* retention closures: src/standalone/Soil/retention_models.jl:30-67 */
enum { CLB_VAN_GENUCHTEN = 0, CLB_BROOKS_COREY = 1 };
/* boundary conditions handled inside the implicit stage
 * (src/standalone/Soil/boundary_conditions.jl:227-411); every other BC type is
 * a flux computed by the host's explicit stage and passed in as a value. */
enum { CLB_TOP_FLUX = 0, CLB_TOP_MOISTURE_STATE = 1 };
enum { CLB_BOT_FLUX = 0, CLB_BOT_FREE_DRAINAGE = 1, CLB_BOT_MOISTURE_STATE = 2 };
/* where a caller pointer lives */
enum { CLB_HOST = 0, CLB_DEVICE = 1 };
/* closure arithmetic: FAST shares log(S) between the powers (exp/log form,
 * <= ~1e-14 relative from the pow form); LIBM evaluates the reference's pow
 * expressions literally with CUDA libm. */
enum { CLB_MATH_FAST = 0, CLB_MATH_LIBM = 1 };
/* kernel variant of the fused step (0 = let the library choose) */
enum { CLB_VARIANT_AUTO = 0,
       CLB_VARIANT_REGISTER_COLUMN = 1, /* one thread per column, N = 15 in registers          */
       CLB_VARIANT_GENERIC = 2,         /* one thread per column, any N, scratch in HBM/L2      */
       CLB_VARIANT_LANE_PER_CELL = 3,   /* one lane per cell, shuffle stencil + cyclic reduction, N <= 31 */
       CLB_VARIANT_LANE_QUAD = 4,       /* four lanes per column (4 cells each), twisted Thomas across the lanes,
                                           TMA-staged stage constants in shared memory; N = 15 or 16,
                                           CLB_MATH_FAST, flux BCs, column-fastest mirrors               */
       CLB_VARIANT_LANE_QUAD_PIPELINED = 5, /* the same as a persistent kernel whose warps prefetch their next
                                           tile (double-buffered shared memory)                          */
       CLB_VARIANT_LANE_OCTET = 6       /* eight lanes per column, Q cells each: 15 <= N <= 64 (Q = ceil(N / 8);
                                           N = 15 / 16 compiled in, the others read at run time) on column-fastest
                                           mirrors, N = 50 (Q = 7) on either layout; otherwise as the quad */ };
/* layout of the library's per-cell mirrors (0 = let the library choose) */
enum { CLB_LAYOUT_AUTO = 0,
       CLB_LAYOUT_COLUMN_FASTEST = 1,   /* element (i, c) at i*ld + c                              */
       CLB_LAYOUT_LEVEL_FASTEST = 2     /* element (i, c) at c*N + i: the reference's own layout    */ };
/* clb_set_option */
enum { CLB_OPT_OUT_OF_PLACE = 1,        /* fused stage reads Y (= temp) and writes the U fields */
       CLB_OPT_CO2_TOP_STATE = 2,       /* SoilCO2Model: the top BC of CO2 is AtmosCO2StateBC (value CLB_F_CO2_C_ATM) */
       CLB_OPT_O2_TOP_STATE = 3,        /* ... of O2 is AtmosO2StateBC (value CLB_F_O2_C_ATM); 0 = flux values */
       CLB_OPT_HOST_ROUTE = 4,          /* clb_implicit_step_host / clb_soil_step_host: 0 = the library's choice (pinned
                                           caller arrays: zero-copy kernels or the copy engines, whichever this host
                                           favours -- measured once per device; pageable ones: staged copies),
                                           1 = copy engines + device staging always, 2 = field by field (clb_set_field /
                                           the step / clb_get_field), 3 = zero-copy kernels for pinned arrays */
       CLB_OPT_HOST_CHUNKS = 5,         /* column chunks of the pipelined host route (0 = default, 4) */
       CLB_OPT_TILE_BOXES = 6,          /* lane kernels: 0 = two TMA boxes of the arena per tile where the mirrors are
                                           equally spaced, 1 = one box per field always (same results, bit for bit) */
       CLB_OPT_EXPLICIT_KERNEL = 7,     /* explicit stage, CLB_MATH_FAST: 0 = warp-uniform control flow (default),
                                           1 = the reference's case distinctions cell by cell (comparison) */
       /* clb_soil_step / clb_soil_step_host: what the explicit stage does about the boundary fluxes */
       CLB_OPT_RUNOFF_MODEL = 8,        /* CLB_RUNOFF_*: how the liquid influx CLB_F_PRECIP becomes the infiltration
                                           (default CLB_RUNOFF_TOPMODEL) */
       CLB_OPT_TOP_ATMOS_DRIVEN = 9,    /* 1: AtmosDrivenFluxBC -- top_bc = (infiltration + vapor_flux_liq, R_n + lhf +
                                           shf + infiltration * e_liq(T_air)); 0 (default): top_bc.water =
                                           infiltration, top_bc.heat as the caller set it */
       CLB_OPT_BOTTOM_EWFD = 10 };      /* 1: EnergyWaterFreeDrainage bottom -- bottom_bc = (-K_1, -K_1 e_liq(T_1));
                                           0 (default): the bottom fluxes as the caller set them */
/* the runoff models of Runoff/Runoff.jl */
enum { CLB_RUNOFF_NONE = 0,             /* NoRunoff :51-75: infiltration = influx */
       CLB_RUNOFF_SURFACE = 1,          /* SurfaceRunoff :78-154: saturation excess at the top cell */
       CLB_RUNOFF_TOPMODEL = 2 };       /* TOPMODELRunoff :157-283 */

/* Field ids.  "cell" fields are N x ncol, "col" fields are ncol. */
typedef enum {
    /* ---- time-invariant parameters (RichardsParameters rre.jl:22-47,
     *      EnergyHydrologyParameters energy_hydrology.jl:60-170) -- cell */
    CLB_F_NU = 0, CLB_F_THETA_R, CLB_F_K_SAT, CLB_F_S_S,
    CLB_F_HCM_A,            /* vG alpha | BC c      */
    CLB_F_HCM_B,            /* vG n     | BC psi_b  */
    CLB_F_HCM_M,            /* vG m     | unused    */
    CLB_F_RHO_C_DS,         /* EH: dry-soil volumetric heat capacity */
    /* ---- lagged cache written by the host's explicit stage -- cell */
    CLB_F_K_LAG, CLB_F_KAPPA_LAG, CLB_F_THETA_L_LAG,   /* EH: p.soil.{K,kappa,theta_l} */
    CLB_F_IS_SATURATED,     /* TOPMODEL source mask, Runoff/Runoff.jl:321-359 */
    /* ---- prognostic state Y -- cell */
    CLB_F_Y_THETA_L, CLB_F_Y_RHO_E_INT, CLB_F_Y_THETA_I,
    /* ---- implicit cache p.soil.{K,psi,T} -- cell */
    CLB_F_P_K, CLB_F_P_PSI, CLB_F_P_T,
    /* ---- implicit tendency dY -- cell */
    CLB_F_DY_THETA_L, CLB_F_DY_RHO_E_INT, CLB_F_DY_THETA_I,
    /* ---- Jacobian blocks, TridiagonalMatrixRow (lower, diag, upper) -- cell */
    CLB_F_W11_LO, CLB_F_W11_DI, CLB_F_W11_UP,   /* (theta_l, theta_l)      */
    CLB_F_W21_LO, CLB_F_W21_DI, CLB_F_W21_UP,   /* (rho_e_int, theta_l) EH */
    CLB_F_W22_LO, CLB_F_W22_DI, CLB_F_W22_UP,   /* (rho_e_int, rho_e_int)  */
    /* ---- linear solve right-hand side b and solution x -- cell */
    CLB_F_B_THETA_L, CLB_F_B_RHO_E_INT, CLB_F_B_THETA_I,
    CLB_F_X_THETA_L, CLB_F_X_RHO_E_INT, CLB_F_X_THETA_I,
    /* ---- new state of the fused stage in out-of-place mode -- cell */
    CLB_F_U_THETA_L, CLB_F_U_RHO_E_INT,
    /* ---- explicit stage of EnergyHydrology (energy_hydrology.jl:722-906) -- cell
     *      parameters only update_aux! reads (EnergyHydrologyParameters :60-170) */
    CLB_F_KAPPA_DRY, CLB_F_KAPPA_SAT_UNFROZEN, CLB_F_KAPPA_SAT_FROZEN,
    CLB_F_NU_SS_OM, CLB_F_NU_SS_QUARTZ, CLB_F_NU_SS_GRAVEL,
    CLB_F_P_TF_DEPRESSED,                    /* p.soil.Tf_depressed */
    CLB_F_DYE_THETA_L, CLB_F_DYE_THETA_I,    /* explicit tendency the PhaseChange source adds into */
    /* ---- SoilCO2Model implicit diffusion (Biogeochemistry.jl:320-413, 1119-1195) -- cell */
    CLB_F_CO2_Y, CLB_F_O2_Y,                 /* Y.soilco2.{CO2, O2} */
    CLB_F_CO2_D, CLB_F_O2_D,                 /* lagged p.soilco2.{D, D_o2} */
    CLB_F_CO2_THETA_EFF, CLB_F_O2_THETA_EFF, /* lagged p.soilco2.{theta_eff, theta_eff_o2} */
    CLB_F_CO2_DY, CLB_F_O2_DY,               /* implicit tendency */
    CLB_F_CO2_W_LO, CLB_F_CO2_W_DI, CLB_F_CO2_W_UP,   /* (CO2, CO2) Jacobian rows */
    CLB_F_O2_W_LO, CLB_F_O2_W_DI, CLB_F_O2_W_UP,      /* (O2, O2) Jacobian rows  */
    CLB_F_CO2_B, CLB_F_O2_B, CLB_F_CO2_X, CLB_F_O2_X, /* right-hand side / solution of the (CO2, CO2), (O2, O2) blocks */
    CLB_F_NUM_CELL,
    /* ---- per-column fields */
    CLB_F_R_SS = CLB_F_NUM_CELL, CLB_F_R_ESS, CLB_F_H_GRAD,     /* lagged TOPMODEL */
    CLB_F_THETA_BC_TOP, CLB_F_THETA_BC_BOT,                     /* MoistureStateBC values */
    CLB_F_TOP_BC_W, CLB_F_BOT_BC_W, CLB_F_TOP_BC_H, CLB_F_BOT_BC_H, /* p.soil.top_bc/bottom_bc */
    CLB_F_DFLUXBCDY, CLB_F_TOTAL_WATER,
    CLB_F_Y_INTF_W, CLB_F_Y_INTF_E,          /* Y.soil.∫F_vol_liq_water_dt, ∫F_e_dt */
    CLB_F_DY_INTF_W, CLB_F_DY_INTF_E,
    CLB_F_B_INTF_W, CLB_F_B_INTF_E, CLB_F_X_INTF_W, CLB_F_X_INTF_E,
    CLB_F_AREA_WEIGHT,                        /* weights of the global balance sums */
    CLB_F_U_INTF_W, CLB_F_U_INTF_E,           /* out-of-place new flux integrals */
    CLB_F_TOTAL_ENERGY,                       /* p.soil.total_energy (explicit update_aux!) */
    CLB_F_F_MAX, CLB_F_PRECIP,                /* TOPMODELRunoff.f_max; the liquid water input (precipitation + melt, m/s, negative down) */
    CLB_F_INFILTRATION, CLB_F_R_S,            /* p.soil.infiltration, p.soil.R_s (clb_update_runoff) */
    CLB_F_CO2_TOP_BC, CLB_F_CO2_BOT_BC, CLB_F_O2_TOP_BC, CLB_F_O2_BOT_BC,   /* p.soilco2.{top,bottom}_bc(_o2) */
    CLB_F_CO2_C_ATM, CLB_F_O2_C_ATM,          /* air-equivalent concentration of the atmosphere (state BC value) */
    CLB_F_CO2_DFLUXBCDY, CLB_F_O2_DFLUXBCDY,  /* p.soilco2.dfluxBCdY(_o2) */
    CLB_F_SFC_W_DI, CLB_F_SFC_B, CLB_F_SFC_X, /* one surface (PointSpace) variable of an integrated model: its
                                                 DiagonalMatrixRow Jacobian block, right-hand side, solution */
    /* the host model's surface fluxes for the atmosphere-driven top boundary (AtmosDrivenFluxBC,
     * boundary_conditions.jl:901-936): p.soil.turbulent_fluxes.{vapor_flux_liq, lhf, shf}, p.soil.R_n, p.drivers.T */
    CLB_F_VAPOR_FLUX_LIQ, CLB_F_LHF, CLB_F_SHF, CLB_F_R_N, CLB_F_T_AIR,
    CLB_F_NUM
} clb_field;

typedef struct {
    int32_t abi_version;          /* = CLB_ABI_VERSION */
    int32_t model;                /* CLB_RICHARDS | CLB_ENERGY_HYDROLOGY */
    int32_t closure;              /* CLB_VAN_GENUCHTEN | CLB_BROOKS_COREY */
    int32_t top_bc, bottom_bc;
    int32_t has_topmodel_source;  /* implicit TOPMODELSubsurfaceRunoff source present */
    int32_t n_levels;             /* N */
    int32_t device;               /* CUDA device ordinal */
    int64_t n_columns;            /* active columns held by this handle (this rank's shard) */
    void *stream;                 /* cudaStream_t; NULL = legacy default stream */
    int32_t math_mode;            /* CLB_MATH_FAST | CLB_MATH_LIBM */
    int32_t kernel_variant;       /* CLB_VARIANT_* */
    /* LandParameters constants (src/shared_utilities/Parameters.jl:11-58) */
    double rho_l, rho_i, cp_l, cp_i, T_ref, LH_f0;
    int32_t layout;               /* CLB_LAYOUT_* */
    int32_t reserved;
} clb_config;

typedef struct {
    int32_t iterations;     /* Newton iterations performed */
    int32_t converged;      /* 1 if the tolerance test passed (always 0 for tol < 0) */
    double dx_norm;         /* ||dx||_2 of the last iteration over all columns of all ranks */
    int64_t nan_count;      /* non-finite entries of the new state (NaNCheckCallback, utils.jl:639-661) */
} clb_stats;

/* ---- lifetime ----------------------------------------------------------- */
int clb_abi_version(void);
/* Per-thread message of the last failure (also valid when handle creation failed). */
const char *clb_last_error(void);
/* Allocates the device mirrors.  Replaces: model construction + initialize(model)
 * for the soil part of Y and p (src/shared_utilities/models.jl:493-499). */
int clb_create(clb_handle *out, const clb_config *cfg);
int clb_destroy(clb_handle h);
/* Blocks until everything enqueued on the handle's stream has finished. */
int clb_sync(clb_handle h);
/* Switch the stream later calls are enqueued on (CUDA.jl task-local stream). */
int clb_set_stream(clb_handle h, void *stream);
/* CLB_OPT_OUT_OF_PLACE: clb_implicit_step keeps Y (ClimaTimeSteppers' `temp`) and
 * writes the new stage value into the CLB_F_U_* fields. */
int clb_set_option(clb_handle h, int32_t option, int64_t value);

/* ---- geometry and masks -------------------------------------------------- */
/* Cell centres z_c[N] and faces z_f[N+1] (host pointers) as ClimaCore produced
 * them (Domains.jl:636-667); thickness and spacings follow Domains.get_dz
 * (Domains.jl:935-949). */
int clb_set_grid(clb_handle h, const double *z_c, const double *z_f);
/* Column j of the handle is column idx[j] of the caller's arrays (land-sea mask
 * compaction; inactive columns are never read or written: test/standalone/Soil/
 * mask_test.jl:53-61).  idx == NULL restores the identity.  Host pointer. */
int clb_set_active_columns(clb_handle h, const int64_t *idx, int64_t n);

/* ---- field transfer ------------------------------------------------------ */
/* Copy a caller array into / out of the library mirror.  Element (i, c) of the
 * caller array is at  ptr[i*stride_level + idx[c]*stride_column]  (strides in
 * elements, >= 0; stride_level is ignored for per-column fields).  `mem` says
 * whether ptr is a host or a device pointer.  A host source has been read completely
 * when clb_set_field returns (it may be reused or freed); a device source is read
 * asynchronously on the handle's stream.  clb_get_field to a host pointer synchronises. */
int clb_set_field(clb_handle h, int32_t field, const double *src, int64_t stride_level,
                  int64_t stride_column, int32_t mem);
int clb_get_field(clb_handle h, int32_t field, double *dst, int64_t stride_level,
                  int64_t stride_column, int32_t mem);
/* Broadcast a scalar parameter (the reference accepts scalars or fields). */
int clb_fill_field(clb_handle h, int32_t field, double value);
/* Device pointer and strides of the library mirror of a field, for callers that
 * want to fill it in place (resident mode). */
int clb_field_device_ptr(clb_handle h, int32_t field, double **ptr, int64_t *stride_level,
                         int64_t *stride_column);

/* Resident state: with the explicit-stage kernels below, a whole soil step can stay in the
 * library's mirrors; the integrator's own vector operations on them are these two.
 *   clb_field_axpy   y <- y + a x   (ARS111's explicit update U0 = u + dt T_exp(u); ClimaTimeSteppers' broadcasts)
 *   clb_field_copy   dst <- src
 * x / y / dst / src: two per-cell or two per-column field ids. */
int clb_field_axpy(clb_handle h, int32_t y, double a, int32_t x);
int clb_field_copy(clb_handle h, int32_t dst, int32_t src);

/* ---- the hooks, fine-grained (parity-checkable per call) ----------------- */
/* update_implicit_cache!(p, Y, t): models.jl:238-246.  Richards: K, psi,
 * total_water and, if the top BC is MoistureStateBC, the boundary fluxes and
 * dfluxBCdY (rre.jl:368-380, 460-468).  EnergyHydrology: T and psi
 * (energy_hydrology.jl:427-455). */
int clb_update_implicit_cache(clb_handle h);
/* Explicit-stage flavour: always evaluates state-type boundary fluxes (rre.jl:111-149). */
int clb_update_boundary_fluxes(clb_handle h);
/* compute_imp_tendency!(dY, Y, p, t): rre.jl:161-203, energy_hydrology.jl:363-425. */
int clb_compute_imp_tendency(clb_handle h);
/* compute_jacobian!(W, Y, p, dtgamma, t): rre.jl:391-458, energy_hydrology.jl:466-576. */
int clb_compute_jacobian(clb_handle h, double dtgamma);
/* ldiv!(x, W, b): BlockDiagonalSolve / BlockLowerTriangularSolve(theta_l),
 * implicit_timestepping.jl:160-171; x = -b for the -I blocks. */
int clb_ldiv(clb_handle h);

/* ---- the explicit stage of EnergyHydrology ------------- */
/* Scalars of EnergyHydrologyParameters (energy_hydrology.jl:150-160: Omega, gamma,
 * gammaT_ref, alpha, beta of the Balland-Arp / impedance / viscosity closures) and the
 * LandParameters constants T_freeze, grav (Parameters.jl:21-22). */
typedef struct {
    double Omega, gamma, gammaT_ref, alpha, beta, T_freeze, grav;
} clb_explicit_params;
int clb_set_explicit_params(clb_handle h, const clb_explicit_params *p);
/* update_aux!(p, Y, t) of EnergyHydrology, energy_hydrology.jl:722-814: from Y.{theta_l, rho_e_int,
 * theta_i} writes p.soil.theta_l -> CLB_F_THETA_L_LAG, kappa -> CLB_F_KAPPA_LAG, K -> CLB_F_K_LAG (the
 * lagged inputs of the implicit stage, produced in place), T -> CLB_F_P_T, psi -> CLB_F_P_PSI,
 * Tf_depressed -> CLB_F_P_TF_DEPRESSED, total_water -> CLB_F_TOTAL_WATER, total_energy ->
 * CLB_F_TOTAL_ENERGY.  RichardsModel's update_aux! is clb_update_implicit_cache (models.jl:207-210). */
int clb_update_aux(clb_handle h);
/* source!(dY, ::PhaseChange, Y, p, model), energy_hydrology.jl:846-906: ADDS -S into
 * CLB_F_DYE_THETA_L and (rho_l / rho_i) S into CLB_F_DYE_THETA_I, S = phase_change_source(p.theta_l,
 * Y.theta_i, p.T, thermal_time(rho_c_s, dz, p.kappa), ...) (soil_heat_parameterizations.jl:35-122). */
int clb_phase_change_source(clb_handle h);
/* Both in one pass over the fields (the source reads the aux values it has just computed): what a
 * resident explicit stage calls. */
int clb_update_aux_and_phase_change(clb_handle h);

/* TOPMODEL runoff of the explicit stage: update_infiltration_water_flux!(p,
 * ::TOPMODELRunoff, input, Y, t, model), src/standalone/Soil/Runoff/Runoff.jl:234-283 (+ :373-434).  From
 * Y, CLB_F_PRECIP (`input`), CLB_F_F_MAX and, for EnergyHydrology, p.soil.{theta_l, T} (clb_update_aux) and
 * the impedance / viscosity scalars of clb_set_explicit_params, writes the lagged inputs of the implicit
 * TOPMODELSubsurfaceRunoff source -- CLB_F_IS_SATURATED, CLB_F_H_GRAD, CLB_F_R_SS, CLB_F_R_ESS -- and
 * CLB_F_INFILTRATION, CLB_F_R_S.  depth: model.domain.depth. */
typedef struct {
    double f_over, R_sb, depth;
} clb_runoff_params;
int clb_set_runoff_params(clb_handle h, const clb_runoff_params *p);
int clb_update_runoff(clb_handle h);

/* The atmosphere-driven top boundary of the explicit stage (second half):
 * soil_boundary_fluxes!(bc::AtmosDrivenFluxBC, Val((:soil,)), model, Y, p, t), boundary_conditions.jl:901-936, after the
 * host model has evaluated turbulent_fluxes! and net_radiation! (SurfaceFluxes.jl and the radiation drivers stay in
 * Julia) and uploaded CLB_F_VAPOR_FLUX_LIQ, CLB_F_LHF, CLB_F_SHF, CLB_F_R_N, CLB_F_T_AIR and the liquid influx
 * CLB_F_PRECIP (compute_liquid_influx :955-961):
 *   update_infiltration_water_flux!(p, runoff, influx, Y, t, model)   runoff_model = CLB_RUNOFF_NONE (Runoff.jl:69-71),
 *       _SURFACE (:129-148: CLB_F_IS_SATURATED = heaviside(theta_l + theta_i - nu), infiltration = (1 - is_saturated at
 *       the top cell) max(i_c, influx), R_s) or _TOPMODEL (clb_update_runoff)
 *   top_bc.water = infiltration + vapor_flux_liq
 *   top_bc.heat  = R_n + lhf + shf + infiltration * volumetric_internal_energy_liq(T_air)        (:988-1002)
 * into CLB_F_TOP_BC_W / _H.  RichardsModel (RichardsAtmosDrivenFluxBC, boundary_flux! :200-213): top_bc = infiltration.
 * EnergyHydrology needs p.soil.{theta_l, T} (clb_update_aux) for the infiltration capacity. */
int clb_update_atmos_driven_fluxes(clb_handle h, int32_t runoff_model);
/* soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...), boundary_conditions.jl:590-608:
 * CLB_F_BOT_BC_W = -K_1, CLB_F_BOT_BC_H = -K_1 * volumetric_internal_energy_liq(T_1) from p.soil.{K, T} of the bottom
 * cell (clb_update_aux). */
int clb_update_energy_water_free_drainage(clb_handle h);

/* ---- SoilCO2Model: implicit CO2 / O2 diffusion ---------- */
/* Two more independent per-column tridiagonals with lagged coefficients, on the same columns and grid as the
 * soil handle (src/standalone/Soil/Biogeochemistry/Biogeochemistry.jl).  With CLB_OPT_*_TOP_STATE the top
 * boundary flux is diffusive_flux(D_N, c_atm, max(C_N / theta_N, 0), dz_top) (:932-957, :1078-1111) and its
 * derivative enters the Jacobian (:1152-1166); otherwise CLB_F_*_TOP_BC holds a flux value.
 *   clb_soilco2_update_boundary_fluxes   make_update_implicit_boundary_fluxes, :320-357
 *   clb_soilco2_compute_imp_tendency     make_compute_imp_tendency, :371-413      -> CLB_F_{CO2,O2}_DY
 *   clb_soilco2_compute_jacobian         make_compute_jacobian, :1119-1195        -> CLB_F_{CO2,O2}_W_*
 *   clb_soilco2_implicit_step            the ARS111 stage (Simulations.jl:127-135): max_iters x (boundary
 *                                        fluxes, Jacobian, tendency, residual, Thomas, update) in one kernel,
 *                                        on CLB_F_{CO2,O2}_Y in place (SOC has no implicit piece, :411) */
int clb_soilco2_update_boundary_fluxes(clb_handle h);
int clb_soilco2_compute_imp_tendency(clb_handle h);
int clb_soilco2_compute_jacobian(clb_handle h, double dtgamma);
int clb_soilco2_implicit_step(clb_handle h, double dtgamma, int32_t max_iters);

/* Diagonal Jacobian blocks: the implicit variables of an integrated model that live on the
 * surface space (canopy.energy.T, src/standalone/Vegetation/canopy_energy.jl:222-250; snow, lake) have
 * DiagonalMatrixRow blocks (implicit_timestepping.jl:117-121), solved entry by entry: x = b / w.  With this the
 * whole ldiv! of SoilCanopyModel / LandModel stays on the device: clb_ldiv for the soil blocks,
 * clb_soilco2_* for CO2 / O2, this for each surface variable (fields CLB_F_SFC_* or any three field ids of
 * one kind). */
int clb_ldiv_diagonal(clb_handle h, int32_t w_field, int32_t b_field, int32_t x_field);

/* `ldiv!(x, W, b)` of an INTEGRATED model's FieldMatrixWithSolver (implicit_timestepping.jl:63-172: the blocks
 * initialize_jacobian builds for the variables present in Y) as ONE launch.  `blocks` says which are present:
 *   CLB_LDIV_SOIL     (theta_l, theta_l) [+ (rho_e_int, theta_l), (rho_e_int, rho_e_int), BlockLowerTriangularSolve] and the
 *                     -I blocks of theta_i and the flux integrals: what clb_ldiv solves (fields CLB_F_W11_* .. _X_*)
 *   CLB_LDIV_SOILCO2  the TridiagonalMatrixRow blocks (soilco2.CO2, soilco2.CO2), (soilco2.O2, soilco2.O2)
 *                     (CLB_F_CO2_W_*, CLB_F_O2_W_* from clb_soilco2_compute_jacobian; CLB_F_CO2_B / _O2_B -> CLB_F_CO2_X / _O2_X)
 *   CLB_LDIV_SURFACE  the DiagonalMatrixRow block of one surface variable, canopy.energy.T (canopy_energy.jl:222-250):
 *                     CLB_F_SFC_X = CLB_F_SFC_B / CLB_F_SFC_W_DI
 * The blocks are independent of each other (BlockDiagonalSolve between models), so the grid's y index runs over them.
 * The remaining explicit variables (-I blocks: x = -b) are a broadcast the caller keeps. */
enum { CLB_LDIV_SOIL = 1, CLB_LDIV_SOILCO2 = 2, CLB_LDIV_SURFACE = 4 };
int clb_ldiv_all(clb_handle h, uint32_t blocks);

/* ---- the fused implicit stage -------------------------------------------- */
/* One implicit ARS111 stage on the resident state Y (in: U = temp, out: new U):
 * cache_imp!, then max_iters x (Wfact, T_imp!, residual, ldiv!, update), all in
 * one kernel when tol < 0 (the reference default: fixed iterations, no
 * convergence checker, Simulations.jl:127-135).  With tol >= 0 the iterations
 * are separate launches that stop -- without a host round trip -- once
 * ||dx||_2 <= tol over all columns of all ranks.  stats may be NULL; when it is
 * not, the call synchronises the stream to fill it.  With a communicator (clb_comm_init)
 * a non-NULL stats and the tol >= 0 path issue an all-reduce: every rank must then make
 * the same choice (stats NULL or not, same tol) in the same call. */
int clb_implicit_step(clb_handle h, double dtgamma, int32_t max_iters, double tol, clb_stats *stats);

/* CLB_VARIANT_* the last clb_implicit_step launched (what CLB_VARIANT_AUTO resolved to; 0 before the first step). */
int clb_last_variant(clb_handle h, int32_t *variant);

/* Host-buffer convenience for the end-to-end path: upload the per-step inputs
 * (state and lagged cache, level-fastest host arrays with column stride N),
 * run the fused stage, download the new state.  n_in / n_out fields. */
int clb_implicit_step_host(clb_handle h, double dtgamma, int32_t max_iters,
                           const int32_t *in_fields, const double *const *in_ptrs, int32_t n_in,
                           const int32_t *out_fields, double *const *out_ptrs, int32_t n_out);

/* A WHOLE soil step of EnergyHydrology from and to host arrays holding the state at t_n -- what
 * ClimaTimeSteppers' step_u! does for the soil with IMEXAlgorithm(ARS111, NewtonsMethod(max_iters))
 * (src/simulations/Simulations.jl:127-135), with every tendency on the device:
 *   update_aux!(p, Y) + source!(dY, ::PhaseChange)   energy_hydrology.jl:722-814, 838-906
 *   update_infiltration_water_flux!(::TOPMODELRunoff)  Runoff/Runoff.jl:234-283 (its infiltration is the top
 *                                                      water flux of the stage; other top fluxes: CLB_F_TOP_BC_*)
 *   U0 = u + dt T_exp(u)                               theta_l, theta_i += dt (PhaseChange source)
 *   the implicit ARS111 stage                          clb_implicit_step(h, dt, max_iters)
 * in_fields: whatever changed on the host since the last call -- normally Y (CLB_F_Y_THETA_L, _Y_RHO_E_INT,
 * _Y_THETA_I, _Y_INTF_W, _Y_INTF_E), CLB_F_PRECIP and the heat / bottom boundary fluxes; the lagged cache (K, kappa,
 * theta_l, is_saturated, R_ss, R_ess, h_grad) is computed on the device and never crosses PCIe.  out_fields: the new
 * state (the Y fields, or the U fields with CLB_OPT_OUT_OF_PLACE).  The columns are cut into chunks whose transfers
 * overlap the neighbouring chunks' kernels: pinned host arrays are read and written in place by the relayout kernels
 * or moved by the copy engines through a device staging area (CLB_OPT_HOST_ROUTE; by default whichever this host
 * favours), pageable ones go through the staging area (same results, bit for bit).  Needs clb_set_explicit_params, clb_set_runoff_params, the explicit-stage
 * parameter fields and CLB_F_F_MAX.  Synchronous. */
int clb_soil_step_host(clb_handle h, double dt, int32_t max_iters,
                       const int32_t *in_fields, const double *const *in_ptrs, int32_t n_in,
                       const int32_t *out_fields, double *const *out_ptrs, int32_t n_out);

/* The same whole soil step on RESIDENT state: the mirrors hold the state at t_n (CLB_F_Y_*), the
 * forcing (CLB_F_PRECIP, CLB_F_TOP_BC_H, CLB_F_BOT_BC_*) and the parameters; on return they hold the state at
 * t_n + dt (in the Y fields, or in the U fields with CLB_OPT_OUT_OF_PLACE -- theta_i always in CLB_F_Y_THETA_I) and the
 * cache of the explicit stage (p.soil.{theta_l, kappa, K, T, psi, Tf_depressed, total_water, total_energy,
 * is_saturated, h_grad, R_ss, R_ess, infiltration, R_s}).  Three launches: the explicit cells, the per-column sweep
 * (runoff + column integrals + explicit update), the fused implicit stage.  Asynchronous on the handle's stream. */
int clb_soil_step(clb_handle h, double dt, int32_t max_iters);

/* ---- diagnostics / reductions -------------------------------------------- */
/* out[c] = sum_i field[i,c] * dz_c[i]  (ClimaCore column_integral_definite!,
 * rre.jl:502-511).  out is a per-column field id. */
int clb_column_integral(clb_handle h, int32_t cell_field, int32_t col_field_out);
/* Weighted global sums for the water / energy balance (definition
 * ext/land_sim_vis/plotting_utils.jl:160-178): out[0] = sum_c w_c * column
 * water, out[1] = sum_c w_c * intF_w, out[2] = sum_c w_c * column energy,
 * out[3] = sum_c w_c * intF_e.  All-reduced over ranks when a communicator is
 * attached.  Synchronous; out is a host pointer to 4 doubles. */
int clb_global_balance(clb_handle h, double *out4);

/* ---- testing ------------------------------------------------------------- */
/* Evaluates one of the library's device math functions (csrc/soil_math.cuh) on host
 * arrays: kind 0 rcp(x), 1 div(x, y), 2 log(x), 3 exp(x), 4 sqrt(x), 5 rcp seed.
 * Used by tests/test_cuda_math.py only. */
int clb_test_math(int32_t kind, const double *x, const double *y, double *out, int64_t n);

/* ---- multi-GPU (one handle per rank; columns are sharded, no halo) -------- */
/* NCCL is loaded at run time (libnccl.so.2).  Rank 0 creates the id, the host
 * side broadcasts the 128 bytes (ClimaComms / torch.distributed), every rank
 * calls clb_comm_init. */
int clb_comm_unique_id(void *id128);
int clb_comm_init(clb_handle h, const void *id128, int32_t n_ranks, int32_t rank);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CLIMALAND_B200_H */
