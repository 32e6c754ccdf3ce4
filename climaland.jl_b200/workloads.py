"""Synthetic inputs of the implicit soil-column path.
Host-side numpy only; shapes are the reference layout: per-cell (ncol, N) level fastest
with level 0 at the bottom, per-column (ncol,).  Parameter distributions are centred on
the reference's fallback means (src/standalone/Soil/spatially_varying_parameters.jl:
295-335), the initial state is a hydrostatic profile above a per-column water table
plus noise (experiments/benchmarks/richards.jl:98-119).  The vertical grid is OUR
geometric stretching between the reference's end thicknesses (Domains.jl:1296-1298:
depth 50 m, dz_tuple = (10, 0.05)); it is not claimed equal to ClimaCore's
GeneralizedExponentialStretching."""
import numpy as np

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))
EARTH = dict(rho_l=1000.0, rho_i=916.7, cp_l=4181.0, cp_i=2100.0, T_ref=273.16,
             LH_f0=2.8344e6 - 2.5008e6)


def stretched_grid(N, depth=50.0, dz_top=0.05):
    """Faces z_f[0..N] from -depth to 0 with geometrically growing thickness downward."""
    lo, hi = 1.0 + 1e-9, 10.0
    for _ in range(200):
        r = 0.5 * (lo + hi)
        s = dz_top * (r ** N - 1.0) / (r - 1.0)
        lo, hi = (r, hi) if s < depth else (lo, r)
    dz = dz_top * r ** np.arange(N)          # top -> bottom
    dz *= depth / dz.sum()
    z_f = np.concatenate([[0.0], -np.cumsum(dz)])[::-1].copy()
    z_f[-1] = 0.0
    return z_f, 0.5 * (z_f[1:] + z_f[:-1])


def vg_theta_from_psi(psi, nu, theta_r, alpha, n, m):
    S = np.where(psi < 0, (1.0 + (alpha * np.abs(psi)) ** n) ** (-m), 1.0)
    return theta_r + S * (nu - theta_r)


def vg_K(theta, nu_eff, theta_r, K_sat, m):
    S = (np.maximum(theta, theta_r + SQRT_EPS) - theta_r) / (np.maximum(nu_eff, theta_r + SQRT_EPS) - theta_r)
    Sc = np.minimum(S, 1.0)
    K = np.sqrt(Sc) * (1.0 - (1.0 - Sc ** (1.0 / m)) ** m) ** 2
    return np.where(S < 1.0, K, 1.0) * K_sat


def make_workload(model, ncol, N=15, seed=0, topmodel=False, ice=True, depth=50.0, dz_top=0.05):
    """model: 'richards' | 'energy_hydrology'.  Returns a dict with the grid, the
    configuration and every field of the implicit path."""
    rng = np.random.default_rng(seed)
    z_f, z_c = stretched_grid(N, depth, dz_top)
    shp = (ncol, N)
    alpha = np.clip(10.0 ** rng.normal(0.14, 0.3, shp), 0.1, 20.0)
    n = np.clip(rng.normal(1.52, 0.2, shp), 1.1, 3.0)
    m = 1.0 - 1.0 / n
    K_sat = np.clip(10.0 ** rng.normal(-5.4, 0.7, shp), SQRT_EPS, 5e-5)
    nu = rng.uniform(0.35, 0.55, shp)
    theta_r = rng.uniform(0.03, 0.12, shp)
    S_s = np.full(shp, 1e-3)
    z_wt = rng.uniform(-40.0, -10.0, (ncol, 1))
    psi = z_wt - z_c[None, :]                       # hydrostatic: psi + z = z_wt
    theta = vg_theta_from_psi(psi, nu, theta_r, alpha, n, m)
    theta = theta * (1.0 + rng.uniform(-0.02, 0.02, shp))
    theta = np.clip(theta, theta_r + 1e-3, nu + 5e-3)
    w = dict(model=model, N=N, ncol=ncol, z_f=z_f, z_c=z_c, topmodel=bool(topmodel),
             nu=nu, theta_r=theta_r, K_sat=K_sat, S_s=S_s, hcm_a=alpha, hcm_b=n, hcm_m=m,
             y_theta_l=theta,
             top_bc_w=-rng.uniform(0.0, 1e-7, ncol), bot_bc_w=np.zeros(ncol),
             y_intf_w=np.zeros(ncol))
    if topmodel:
        w["is_saturated"] = (theta >= nu).astype(np.float64)
        w["r_ss"] = rng.uniform(0.0, 1e-8, ncol)
        w["h_grad"] = rng.uniform(0.0, 5.0, ncol)
    if model == "energy_hydrology":
        E = EARTH
        T = rng.uniform(265.0, 300.0, shp)
        theta_i = np.where(T < 273.0, rng.uniform(0.0, 0.1, shp), 0.0) if ice else np.zeros(shp)
        theta = np.minimum(theta, nu - theta_i + 5e-3)
        theta = np.maximum(theta, theta_r + 1e-3)
        theta_l = np.minimum(nu - theta_i, theta)
        rho_c_ds = 2.0e6 * (1.0 - nu)
        rho_c_s = rho_c_ds + theta_l * E["rho_l"] * E["cp_l"] + theta_i * E["rho_i"] * E["cp_i"]
        rho_e = rho_c_s * (T - E["T_ref"]) - theta_i * E["rho_i"] * E["LH_f0"]
        # lagged cache of the explicit stage: K = impedance * viscosity * vG K (energy_hydrology.jl:745-789)
        f_i = theta_i / (theta_l + theta_i)
        K_lag = 10.0 ** (-7.0 * f_i) * np.exp(2.64e-2 * (T - 288.0)) * vg_K(theta, nu - theta_i, theta_r, K_sat, m)
        w.update(y_theta_l=theta, y_rho_e_int=rho_e, y_theta_i=theta_i, rho_c_ds=rho_c_ds,
                 k_lag=K_lag, kappa_lag=rng.uniform(0.4, 2.6, shp), theta_l_lag=theta_l,
                 top_bc_h=rng.uniform(-50.0, 50.0, ncol), bot_bc_h=np.zeros(ncol),
                 y_intf_e=np.zeros(ncol))
        if topmodel:
            w["is_saturated"] = (theta >= nu - theta_i).astype(np.float64)
            w["r_ess"] = w["r_ss"] * E["rho_l"] * E["cp_l"] * 10.0
    return w


# scalars of EnergyHydrologyParameters pinned by the reference's own test
# (test/standalone/Soil/soil_parameterizations.jl:71-76); T_freeze, grav: ClimaParams defaults
EXPLICIT_SCALARS = dict(Omega=7.0, gamma=2.64e-2, gammaT_ref=288.0, alpha=0.24, beta=18.3, T_freeze=273.15, grav=9.81)


def make_explicit_params(w, seed=0):
    """The six per-cell parameter fields only the explicit stage reads (update_aux!,
    energy_hydrology.jl:722-814), derived as EnergyHydrologyParameters derives them
    (soil_heat_parameterizations.jl:338-395) from synthetic soil composition fractions."""
    rng = np.random.default_rng(10_000 + seed)
    shp = w["nu"].shape
    nu = w["nu"]
    om = rng.uniform(0.0, 0.2, shp)
    quartz = rng.uniform(0.1, 0.5, shp)
    gravel = rng.uniform(0.0, 0.2, shp)
    k_om, k_quartz, k_min, k_ice, k_liq, k_air = 0.25, 8.0, 2.5, 2.21, 0.57, 0.025
    k_solid = k_om ** om * k_quartz ** quartz * k_min ** (1.0 - om - quartz)
    rho_p = 2700.0
    rho_b = (1.0 - nu) * rho_p
    return dict(nu_ss_om=om, nu_ss_quartz=quartz, nu_ss_gravel=gravel,
                kappa_sat_frozen=k_solid ** (1.0 - nu) * k_ice ** nu,
                kappa_sat_unfrozen=k_solid ** (1.0 - nu) * k_liq ** nu,
                kappa_dry=((0.053 * k_solid - k_air) * rho_b + k_air * rho_p) / (rho_p - (1.0 - 0.053) * rho_b))


CELL_PARAMS = ("nu", "theta_r", "K_sat", "S_s", "hcm_a", "hcm_b", "hcm_m", "rho_c_ds", "k_lag", "kappa_lag",
               "theta_l_lag", "is_saturated")
COL_PARAMS = ("r_ss", "r_ess", "h_grad", "theta_bc_top", "theta_bc_bot")
STATE = ("y_theta_l", "y_rho_e_int", "y_theta_i", "y_intf_w", "y_intf_e")
BCS = ("top_bc_w", "bot_bc_w", "top_bc_h", "bot_bc_h")

# algorithmic bytes per column-step of the fused stage (SURVEY 8d / BASELINE.md section 2)
def algorithmic_bytes(model, N, topmodel=False):
    if model == "richards":
        n3_in, n3_out, n2 = 8 + (1 if topmodel else 0), 1, 4 + (2 if topmodel else 0)
    else:
        n3_in, n3_out, n2 = 13 + (1 if topmodel else 0), 2, 8 + (3 if topmodel else 0)
    return 8 * (N * (n3_in + n3_out) + n2)
