"""Host-side mirror of ClimaLand's interface for the implicit soil-column path.

The reference's host language is Julia (no toolchain in this image); the real binding is
julia/ClimaLandB200.jl (`ccall` over the same C ABI).  This module
restates the same surface in Python -- same names, argument meaning and error behaviour,
minus the `!` -- so the parity tests read like the reference's own tests:

    RichardsParameters / EnergyHydrologyParameters   rre.jl:7-47, energy_hydrology.jl:60-170
    vanGenuchten / BrooksCorey                       retention_models.jl:30-67
    MoistureStateBC, WaterFluxBC, FreeDrainage, HeatFluxBC, WaterHeatBC
                                                     boundary_conditions.jl:60-353
    TOPMODELSubsurfaceRunoff                         Runoff/Runoff.jl:174-183
    RichardsModel / EnergyHydrology                  rre.jl:67-109, energy_hydrology.jl:236-300
    initialize                                       shared_utilities/models.jl:493-499
    make_update_implicit_cache / make_compute_imp_tendency / make_compute_jacobian
                                                     models.jl:238-306, implicit_timestepping.jl:25-28
    initialize_jacobian + ldiv                       implicit_timestepping.jl:63-172
    make_update_aux (EnergyHydrology) / PhaseChange + source
                                                     energy_hydrology.jl:722-814, 826-906
    IMEXAlgorithm(ARS111, NewtonsMethod) / LandSimulation / step / solve
                                                     simulations/Simulations.jl:115-332

Y, p and dY are plain namespaces of numpy arrays in the reference layout ((ncol, N), level
fastest, level 0 at the bottom).  Every hook moves its inputs to the device, calls the C ABI
and brings its outputs back: that is the fine-grained, parity-checkable drop-in level.  The
performance level is `FusedSoilNewton` (state resident on the device, one kernel per stage).
All arithmetic of the path happens in libclimaland_b200.so."""
from types import SimpleNamespace

import numpy as np

from . import solver as _s


# ---- retention closures ------------------------------------------------------------------
class vanGenuchten:
    def __init__(self, *, α, n, m=None):
        self.α, self.n = α, n
        self.m = (1 - 1 / np.asarray(n, dtype=np.float64)) if m is None else m
    closure = _s.VAN_GENUCHTEN

    def abm(self):
        return self.α, self.n, self.m


class BrooksCorey:
    def __init__(self, *, ψb, c):
        self.ψb, self.c = ψb, c
    closure = _s.BROOKS_COREY

    def abm(self):
        return self.c, self.ψb, 0.0


# ---- parameters ----------------------------------------------------------------------------
class RichardsParameters:
    def __init__(self, *, hydrology_cm, ν, K_sat, S_s, θ_r):
        self.hydrology_cm, self.ν, self.K_sat, self.S_s, self.θ_r = hydrology_cm, ν, K_sat, S_s, θ_r


class EnergyHydrologyParameters(RichardsParameters):
    """energy_hydrology.jl:60-170.  The implicit path needs ρc_ds only; the explicit stage
    (make_update_aux, PhaseChange) also reads the thermal-conductivity fields and the scalar
    closures' constants, whose defaults are the reference's (test/standalone/Soil/
    soil_parameterizations.jl:71-76); T_freeze / grav belong to earth_param_set in the reference."""
    def __init__(self, *, hydrology_cm, ν, K_sat, S_s, θ_r, ρc_ds, earth_param_set=None, κ_dry=None,
                 κ_sat_unfrozen=None, κ_sat_frozen=None, ν_ss_om=None, ν_ss_quartz=None, ν_ss_gravel=None,
                 α=0.24, β=18.3, γ=2.64e-2, γT_ref=288.0, Ω=7.0, T_freeze=273.15, grav=9.81):
        super().__init__(hydrology_cm=hydrology_cm, ν=ν, K_sat=K_sat, S_s=S_s, θ_r=θ_r)
        self.ρc_ds = ρc_ds
        self.earth_param_set = dict(_s.EARTH if earth_param_set is None else earth_param_set)
        self.κ_dry, self.κ_sat_unfrozen, self.κ_sat_frozen = κ_dry, κ_sat_unfrozen, κ_sat_frozen
        self.ν_ss_om, self.ν_ss_quartz, self.ν_ss_gravel = ν_ss_om, ν_ss_quartz, ν_ss_gravel
        self.α, self.β, self.γ, self.γT_ref, self.Ω, self.T_freeze, self.grav = α, β, γ, γT_ref, Ω, T_freeze, grav

    def has_explicit_fields(self):
        return all(v is not None for v in (self.κ_dry, self.κ_sat_unfrozen, self.κ_sat_frozen, self.ν_ss_om,
                                           self.ν_ss_quartz, self.ν_ss_gravel))


# ---- boundary conditions -------------------------------------------------------------------
class MoistureStateBC:
    """bc(p, t) -> theta_l at the boundary (scalar or per-column array)"""
    def __init__(self, bc):
        self.bc = bc if callable(bc) else (lambda p, t, v=bc: v)


class WaterFluxBC:
    """bc(p, t) -> water flux (m/s, positive upward)"""
    def __init__(self, bc):
        self.bc = bc if callable(bc) else (lambda p, t, v=bc: v)


class HeatFluxBC(WaterFluxBC):
    pass


class FreeDrainage:
    pass


class WaterHeatBC:
    def __init__(self, *, water, heat):
        self.water, self.heat = water, heat


class TOPMODELSubsurfaceRunoff:
    """Implicit source: the lagged p.soil.{R_ss, R_ess, h∇, is_saturated} are inputs."""
    explicit = False


class PhaseChange:
    """PhaseChange source (energy_hydrology.jl:826-829): explicit in every prognostic variable."""
    explicit = True


# ---- domain --------------------------------------------------------------------------------
class Column:
    """`ncol` independent columns sharing one vertical grid (Column / HybridBox / SphericalShell
    all look like this to the implicit path: Domains.jl:23-1446).  z_f: N+1 faces, bottom -> top."""
    def __init__(self, *, zlim=None, nelements=None, z_f=None, z_c=None, ncol=1, active_columns=None):
        if z_f is None:
            z_f = np.linspace(zlim[0], zlim[1], nelements + 1)
        self.z_f = np.ascontiguousarray(z_f, dtype=np.float64)
        self.z_c = 0.5 * (self.z_f[1:] + self.z_f[:-1]) if z_c is None else np.ascontiguousarray(z_c)
        self.N = self.z_f.size - 1
        self.ncol = int(ncol)
        self.active_columns = active_columns   # land-sea mask (indices of the active columns)


# ---- models --------------------------------------------------------------------------------
class _SoilModel:
    kind = None

    def __init__(self, *, parameters, domain, boundary_conditions, sources=(), lateral_flow=False, device=0,
                 math_mode=_s.MATH_FAST, kernel_variant=_s.VARIANT_AUTO, layout=_s.LAYOUT_AUTO):
        assert not lateral_flow, "lateral flow is not part of the implicit column path (rre.jl:106)"
        self.parameters, self.domain, self.sources = parameters, domain, tuple(sources)
        bc = boundary_conditions
        self.boundary_conditions = bc if isinstance(bc, SimpleNamespace) else SimpleNamespace(**bc)
        top, bot = self._water(self.boundary_conditions.top), self._water(self.boundary_conditions.bottom)
        self._top_kind = _s.TOP_MOISTURE_STATE if isinstance(top, MoistureStateBC) else _s.TOP_FLUX
        self._bot_kind = (_s.BOT_MOISTURE_STATE if isinstance(bot, MoistureStateBC)
                          else _s.BOT_FREE_DRAINAGE if isinstance(bot, FreeDrainage) else _s.BOT_FLUX)
        self.has_topmodel = any(isinstance(s, TOPMODELSubsurfaceRunoff) for s in self.sources)
        d = domain
        act = d.active_columns
        n_active = d.ncol if act is None else len(act)
        earth = getattr(parameters, "earth_param_set", None)
        self.solver = _s.SoilColumnSolver(
            model=self.kind, n_columns=n_active, n_columns_total=d.ncol, z_f=d.z_f, z_c=d.z_c,
            closure=parameters.hydrology_cm.closure, top_bc=self._top_kind, bottom_bc=self._bot_kind,
            has_topmodel_source=self.has_topmodel, device=device, math_mode=math_mode,
            kernel_variant=kernel_variant, layout=layout, earth=earth, active_columns=act)
        a, b, m = parameters.hydrology_cm.abm()
        for name, v in (("nu", parameters.ν), ("theta_r", parameters.θ_r), ("K_sat", parameters.K_sat),
                        ("S_s", parameters.S_s), ("hcm_a", a), ("hcm_b", b), ("hcm_m", m)):
            self._set_param(name, v)
        if self.kind == _s.ENERGY_HYDROLOGY:
            self._set_param("rho_c_ds", parameters.ρc_ds)
            if getattr(parameters, "has_explicit_fields", lambda: False)():
                q = parameters
                for name, v in (("kappa_dry", q.κ_dry), ("kappa_sat_unfrozen", q.κ_sat_unfrozen),
                                ("kappa_sat_frozen", q.κ_sat_frozen), ("nu_ss_om", q.ν_ss_om),
                                ("nu_ss_quartz", q.ν_ss_quartz), ("nu_ss_gravel", q.ν_ss_gravel)):
                    self._set_param(name, v)
                self.solver.set_explicit_params(Omega=q.Ω, gamma=q.γ, gammaT_ref=q.γT_ref, alpha=q.α, beta=q.β,
                                                T_freeze=q.T_freeze, grav=q.grav)

    @staticmethod
    def _water(bc):
        return bc.water if isinstance(bc, WaterHeatBC) else bc

    @staticmethod
    def _heat(bc):
        return bc.heat if isinstance(bc, WaterHeatBC) else None

    def _set_param(self, name, v):
        v = np.asarray(v, dtype=np.float64)
        self.solver.set(name, float(v) if v.ndim == 0 else v)

    # prognostic / auxiliary variable lists (rre.jl:282-355, energy_hydrology.jl:624-710)
    def _cell(self):
        return np.zeros((self.domain.ncol, self.domain.N))

    def _col(self):
        return np.zeros(self.domain.ncol)


class RichardsModel(_SoilModel):
    kind = _s.RICHARDS


class EnergyHydrology(_SoilModel):
    kind = _s.ENERGY_HYDROLOGY


def initialize(model):
    """-> Y, p, coords (models.jl:493-499).  Only the soil variables the implicit path touches."""
    eh = model.kind == _s.ENERGY_HYDROLOGY
    soilY = SimpleNamespace(ϑ_l=model._cell(), ᶠF_vol_liq_water_dt=model._col())
    soilp = SimpleNamespace(K=model._cell(), ψ=model._cell(), total_water=model._col(),
                            top_bc=model._col(), bottom_bc=model._col())
    if model._top_kind == _s.TOP_MOISTURE_STATE and not eh:
        soilp.dfluxBCdY = model._col()
    if eh:
        soilY.ρe_int, soilY.θ_i, soilY.ᶠF_e_dt = model._cell(), model._cell(), model._col()
        soilp.T, soilp.κ, soilp.θ_l = model._cell(), model._cell(), model._cell()
        soilp.top_bc = SimpleNamespace(water=model._col(), heat=model._col())
        soilp.bottom_bc = SimpleNamespace(water=model._col(), heat=model._col())
    if model.has_topmodel:
        soilp.R_ss, soilp.R_ess, soilp.h_grad, soilp.is_saturated = (model._col(), model._col(), model._col(),
                                                                      model._cell())
    coords = SimpleNamespace(subsurface=SimpleNamespace(z=np.broadcast_to(model.domain.z_c, soilY.ϑ_l.shape)))
    return SimpleNamespace(soil=soilY), SimpleNamespace(soil=soilp), coords


# ---- host <-> device plumbing of the hooks --------------------------------------------------
def _push_state(model, Y):
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    s.set("y_theta_l", Y.soil.ϑ_l)
    s.set("y_intf_w", Y.soil.ᶠF_vol_liq_water_dt)
    if eh:
        s.set("y_rho_e_int", Y.soil.ρe_int)
        s.set("y_theta_i", Y.soil.θ_i)
        s.set("y_intf_e", Y.soil.ᶠF_e_dt)


def _pull_state(model, Y, prefix="y"):
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    s.get(f"{prefix}_theta_l", Y.soil.ϑ_l)
    s.get(f"{prefix}_intf_w", Y.soil.ᶠF_vol_liq_water_dt)
    if eh:
        s.get(f"{prefix}_rho_e_int", Y.soil.ρe_int)
        if prefix != "u":
            s.get(f"{prefix}_theta_i", Y.soil.θ_i)
        s.get(f"{prefix}_intf_e", Y.soil.ᶠF_e_dt)


def _push_lagged(model, p, t):
    """Inputs the host's explicit stage owns: flux-type boundary values, state-type boundary
    theta, lagged K / kappa / theta_l (EnergyHydrology) and the TOPMODEL fields."""
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    bcs = model.boundary_conditions
    top, bot = model._water(bcs.top), model._water(bcs.bottom)
    ptw = p.soil.top_bc.water if eh else p.soil.top_bc
    pbw = p.soil.bottom_bc.water if eh else p.soil.bottom_bc
    if isinstance(top, MoistureStateBC):
        s.set("theta_bc_top", np.broadcast_to(np.asarray(top.bc(p, t), dtype=np.float64), ptw.shape))
    elif isinstance(top, WaterFluxBC):
        ptw[...] = top.bc(p, t)
    if isinstance(bot, MoistureStateBC):
        s.set("theta_bc_bot", np.broadcast_to(np.asarray(bot.bc(p, t), dtype=np.float64), pbw.shape))
    elif isinstance(bot, WaterFluxBC):
        pbw[...] = bot.bc(p, t)
    s.set("top_bc_w", ptw)
    s.set("bot_bc_w", pbw)
    if eh:
        for side, name in ((bcs.top, "top"), (bcs.bottom, "bot")):
            heat = model._heat(side)
            arr = getattr(p.soil, "top_bc" if name == "top" else "bottom_bc").heat
            if isinstance(heat, HeatFluxBC):
                arr[...] = heat.bc(p, t)
            s.set(f"{name}_bc_h", arr)
        s.set("k_lag", p.soil.K)
        s.set("kappa_lag", p.soil.κ)
        s.set("theta_l_lag", p.soil.θ_l)
    if model.has_topmodel:
        s.set("r_ss", p.soil.R_ss)
        s.set("h_grad", p.soil.h_grad)
        s.set("is_saturated", p.soil.is_saturated)
        if eh:
            s.set("r_ess", p.soil.R_ess)


def _pull_cache(model, p):
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    s.get("p_psi", p.soil.ψ)
    if eh:
        s.get("p_t", p.soil.T)
    else:
        s.get("p_k", p.soil.K)
        s.get("total_water", p.soil.total_water)
        if model._top_kind == _s.TOP_MOISTURE_STATE:
            s.get("top_bc_w", p.soil.top_bc)
            s.get("bot_bc_w", p.soil.bottom_bc)
            s.get("dfluxbcdy", p.soil.dfluxBCdY)


def _push_cache(model, p):
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    s.set("p_psi", p.soil.ψ)
    if eh:
        s.set("p_t", p.soil.T)
    else:
        s.set("p_k", p.soil.K)
        if model._top_kind == _s.TOP_MOISTURE_STATE:
            s.set("dfluxbcdy", p.soil.dfluxBCdY)


def make_update_implicit_cache(model):
    """update_implicit_cache!(p, Y, t): models.jl:238-246"""
    def update_implicit_cache(p, Y, t):
        _push_state(model, Y)
        _push_lagged(model, p, t)
        model.solver.update_implicit_cache()
        _pull_cache(model, p)
    return update_implicit_cache


def make_update_aux(model):
    """update_aux!(p, Y, t).  EnergyHydrology: energy_hydrology.jl:722-814 (θ_l, κ, T, K, ψ, Tf_depressed,
    total_water, total_energy).  RichardsModel: rre.jl:368-380, which is also its implicit cache update
    (models.jl:207-210)."""
    if model.kind == _s.RICHARDS:
        return make_update_implicit_cache(model)

    def update_aux(p, Y, t):
        _push_state(model, Y)
        s = model.solver
        s.update_aux()
        for dev, name in (("theta_l_lag", "θ_l"), ("kappa_lag", "κ"), ("k_lag", "K"), ("p_t", "T"), ("p_psi", "ψ"),
                          ("p_tf_depressed", "Tf_depressed"), ("total_water", "total_water"),
                          ("total_energy", "total_energy")):
            if not hasattr(p.soil, name):
                setattr(p.soil, name, model._col() if name.startswith("total") else model._cell())
            s.get(dev, getattr(p.soil, name))
    return update_aux


def make_phase_change_source(model):
    """source!(dY, src::PhaseChange, Y, p, model): energy_hydrology.jl:846-906; adds to dY.soil.ϑ_l, θ_i."""
    def source(dY, src, Y, p, model_=None):
        assert isinstance(src, PhaseChange)
        s = model.solver
        _push_state(model, Y)
        s.set("theta_l_lag", p.soil.θ_l)
        s.set("kappa_lag", p.soil.κ)
        s.set("p_t", p.soil.T)
        s.set("dye_theta_l", dY.soil.ϑ_l)
        s.set("dye_theta_i", dY.soil.θ_i)
        s.phase_change_source()
        s.get("dye_theta_l", dY.soil.ϑ_l)
        s.get("dye_theta_i", dY.soil.θ_i)
    return source


def make_update_boundary_fluxes(model):
    """Explicit-stage flavour (rre.jl:111-149): always evaluates state-type boundary fluxes."""
    def update_boundary_fluxes(p, Y, t):
        _push_state(model, Y)
        _push_lagged(model, p, t)
        _push_cache(model, p)
        model.solver.update_boundary_fluxes()
        eh = model.kind == _s.ENERGY_HYDROLOGY
        model.solver.get("top_bc_w", p.soil.top_bc.water if eh else p.soil.top_bc)
        model.solver.get("bot_bc_w", p.soil.bottom_bc.water if eh else p.soil.bottom_bc)
        if not eh and model._top_kind == _s.TOP_MOISTURE_STATE:
            model.solver.get("dfluxbcdy", p.soil.dfluxBCdY)
    return update_boundary_fluxes


def make_compute_imp_tendency(model):
    """compute_imp_tendency!(dY, Y, p, t): rre.jl:161-203, energy_hydrology.jl:363-425"""
    def compute_imp_tendency(dY, Y, p, t):
        _push_state(model, Y)
        _push_lagged(model, p, t)
        _push_cache(model, p)
        s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
        s.compute_imp_tendency()
        s.get("dy_theta_l", dY.soil.ϑ_l)
        s.get("dy_intf_w", dY.soil.ᶠF_vol_liq_water_dt)
        if eh:
            s.get("dy_rho_e_int", dY.soil.ρe_int)
            s.get("dy_theta_i", dY.soil.θ_i)
            s.get("dy_intf_e", dY.soil.ᶠF_e_dt)
    return compute_imp_tendency


class B200SoilJacobian:
    """jac_prototype of the drop-in (initialize_jacobian, implicit_timestepping.jl:63-172): the
    tridiagonal blocks live on the device; `matrix[(row, col)]` downloads a block as
    (lower, diag, upper) arrays for inspection, as `jacobian.matrix[@name(soil.ϑ_l), ...]` does."""
    def __init__(self, model):
        self.model = model
        eh = model.kind == _s.ENERGY_HYDROLOGY
        # solver_algorithm as the reference picks it (implicit_timestepping.jl:160-171)
        self.solver_algorithm = "BlockLowerTriangularSolve(soil.ϑ_l)" if eh else "BlockDiagonalSolve"
        self.keys = [("soil.ϑ_l", "soil.ϑ_l")] + ([("soil.ρe_int", "soil.ϑ_l"), ("soil.ρe_int", "soil.ρe_int")]
                                                   if eh else [])
        self._blocks = {("soil.ϑ_l", "soil.ϑ_l"): "w11", ("soil.ρe_int", "soil.ϑ_l"): "w21",
                        ("soil.ρe_int", "soil.ρe_int"): "w22"}

    def block(self, key):
        b = self._blocks[key]
        s = self.model.solver
        return tuple(s.get(f"{b}_{d}") for d in ("lo", "di", "up"))


def initialize_jacobian(model):
    return B200SoilJacobian(model)


def make_compute_jacobian(model):
    """compute_jacobian!(jacobian, Y, p, dtγ, t): rre.jl:391-458, energy_hydrology.jl:466-576"""
    def compute_jacobian(jacobian, Y, p, dtγ, t):
        assert isinstance(jacobian, B200SoilJacobian) and jacobian.model is model
        _push_state(model, Y)
        _push_lagged(model, p, t)
        _push_cache(model, p)
        model.solver.compute_jacobian(float(dtγ))
    return compute_jacobian


def ldiv(x, jacobian, b):
    """ldiv!(x, W, b) -> MatrixFields.field_matrix_solve! (implicit_timestepping.jl:160-171)"""
    model = jacobian.model
    s, eh = model.solver, model.kind == _s.ENERGY_HYDROLOGY
    s.set("b_theta_l", b.soil.ϑ_l)
    s.set("b_intf_w", b.soil.ᶠF_vol_liq_water_dt)
    if eh:
        s.set("b_rho_e_int", b.soil.ρe_int)
        s.set("b_theta_i", b.soil.θ_i)
        s.set("b_intf_e", b.soil.ᶠF_e_dt)
    s.ldiv()
    s.get("x_theta_l", x.soil.ϑ_l)
    s.get("x_intf_w", x.soil.ᶠF_vol_liq_water_dt)
    if eh:
        s.get("x_rho_e_int", x.soil.ρe_int)
        s.get("x_theta_i", x.soil.θ_i)
        s.get("x_intf_e", x.soil.ᶠF_e_dt)


# ---- time stepping (ClimaTimeSteppers surface used by Simulations.jl:127-135) -----------------
class NewtonsMethod:
    """max_iters Newton iterations, Jacobian updated every iteration
    (update_j = UpdateEvery(NewNewtonIteration)); tol = None -> no convergence checker."""
    def __init__(self, max_iters=3, tol=None):
        self.max_iters, self.tol = max_iters, tol


class FusedSoilNewton(NewtonsMethod):
    """The performance drop-in: the whole Newton loop of a stage is ONE clb_implicit_step call."""


class IMEXAlgorithm:
    def __init__(self, tableau="ARS111", newtons_method=None):
        assert tableau == "ARS111", "only the reference default ARS111 is mirrored"
        self.newtons_method = newtons_method or NewtonsMethod(max_iters=3)


def _copy_state(Y):
    return SimpleNamespace(soil=SimpleNamespace(**{k: np.array(v, copy=True) for k, v in vars(Y.soil).items()}))


def _axpy(dst, a, src):
    for k, v in vars(src.soil).items():
        getattr(dst.soil, k)[...] += a * v


class LandSimulation:
    """LandSimulation(t0, tf, Δt, model; timestepper, set_ic!, exp_tendency!) (Simulations.jl:115-245).
    exp_tendency(dY, Y, p, t) is the host's explicit stage (update_cache! + explicit sources);
    it defaults to zero tendency, which is RichardsModel's (no lateral flow, no explicit source)."""
    def __init__(self, t0, tf, Δt, model, *, timestepper=None, set_ic=None, exp_tendency=None):
        self.t, self.tf, self.Δt, self.model = float(t0), float(tf), float(Δt), model
        self.timestepper = timestepper or IMEXAlgorithm()
        self.Y, self.p, self.coords = initialize(model)
        if set_ic is not None:
            set_ic(self.Y, self.p, t0, model)
        self.exp_tendency = exp_tendency
        self.imp_tendency = make_compute_imp_tendency(model)
        self.jacobian = make_compute_jacobian(model)
        self.cache_imp = make_update_implicit_cache(model)
        self.jac_prototype = initialize_jacobian(model)
        self.stats = None

    def step(self):
        """One ARS111 step: forward-Euler explicit stage, backward-Euler implicit
        stage by Newton's method, then u = u + dt*T_exp + dt*T_imp."""
        Y, p, dt, t, model = self.Y, self.p, self.Δt, self.t, self.model
        nm = self.timestepper.newtons_method
        T_exp = _copy_state(Y)
        for v in vars(T_exp.soil).values():
            v[...] = 0.0
        if self.exp_tendency is not None:
            self.exp_tendency(T_exp, Y, p, t)
        U = _copy_state(Y)
        _axpy(U, dt, T_exp)
        temp = _copy_state(U)
        tn = t + dt
        if isinstance(nm, FusedSoilNewton):
            _push_state(model, U)
            _push_lagged(model, p, tn)
            self.stats = model.solver.implicit_step(dt, nm.max_iters, -1.0 if nm.tol is None else nm.tol,
                                                    want_stats=True)
            _pull_state(model, U)
        else:
            self.cache_imp(p, U, tn)
            f, dx = _copy_state(U), _copy_state(U)
            for n in range(1, nm.max_iters + 1):
                self.jacobian(self.jac_prototype, U, p, dt, tn)
                self.imp_tendency(f, U, p, tn)
                for k in vars(f.soil):
                    getattr(f.soil, k)[...] = getattr(temp.soil, k) + dt * getattr(f.soil, k) - getattr(U.soil, k)
                ldiv(dx, self.jac_prototype, f)
                _axpy(U, -1.0, dx)
                if nm.tol is not None:
                    nrm = np.sqrt(sum(float(np.sum(v * v)) for v in vars(dx.soil).values()))
                    if nrm <= nm.tol:
                        break
                if n < nm.max_iters:
                    self.cache_imp(p, U, tn)
        # u = u + dt*T_exp + dt*T_imp with T_imp = (U - temp)/dt
        for k in vars(Y.soil):
            T_imp = (getattr(U.soil, k) - getattr(temp.soil, k)) / dt
            getattr(Y.soil, k)[...] = getattr(Y.soil, k) + dt * getattr(T_exp.soil, k) + dt * T_imp
        self.t = tn
        return self

    def solve(self):
        while self.t < self.tf - 1e-9 * self.Δt:
            self.step()
        return self
