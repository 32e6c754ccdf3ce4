"""ctypes front end of the CPU oracle (oracle/soil_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs; never by the product
package.  Arrays are numpy float64 in the reference's CPU layout: per-cell
fields are C-contiguous ``(ncol, N)`` (level fastest, level 0 = bottom),
per-column fields are ``(ncol,)``.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

RICHARDS, ENERGY_HYDROLOGY = 0, 1
VAN_GENUCHTEN, BROOKS_COREY = 0, 1
TOP_FLUX, TOP_MOISTURE_STATE = 0, 1
BOT_FLUX, BOT_FREE_DRAINAGE, BOT_MOISTURE_STATE = 0, 1, 2

# LandParameters constants (ClimaParams 1.1.4 defaults; passed in as numbers,
# src/shared_utilities/Parameters.jl:86-105)
EARTH = dict(rho_l=1000.0, rho_i=916.7, cp_l=4181.0, cp_i=2100.0, T_ref=273.16,
             LH_f0=2.8344e6 - 2.5008e6)

_dp = C.POINTER(C.c_double)


class _Problem(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("closure", C.c_int32), ("top_bc", C.c_int32),
        ("bottom_bc", C.c_int32), ("has_topmodel_source", C.c_int32),
        ("N", C.c_int32), ("ncol", C.c_int64), ("nthreads", C.c_int32),
        ("z_c", _dp), ("z_f", _dp),
        ("nu", _dp), ("theta_r", _dp), ("K_sat", _dp), ("S_s", _dp),
        ("hcm_a", _dp), ("hcm_b", _dp), ("hcm_m", _dp),
        ("rho_c_ds", _dp), ("K_lag", _dp), ("kappa_lag", _dp), ("theta_l_lag", _dp),
        ("is_saturated", _dp), ("R_ss", _dp), ("R_ess", _dp), ("h_grad", _dp),
        ("theta_bc_top", _dp), ("theta_bc_bot", _dp),
        ("rho_l", C.c_double), ("rho_i", C.c_double), ("cp_l", C.c_double),
        ("cp_i", C.c_double), ("T_ref", C.c_double), ("LH_f0", C.c_double),
    ]


class _State(C.Structure):
    _fields_ = [("theta_l", _dp), ("rho_e_int", _dp), ("theta_i", _dp),
                ("intF_w", _dp), ("intF_e", _dp)]


class _Cache(C.Structure):
    _fields_ = [("K", _dp), ("psi", _dp), ("T", _dp), ("top_bc_w", _dp), ("bot_bc_w", _dp),
                ("top_bc_h", _dp), ("bot_bc_h", _dp), ("dfluxBCdY", _dp), ("total_water", _dp)]


class _Jacobian(C.Structure):
    _fields_ = [(n, _dp) for n in ("w11_lo", "w11_di", "w11_up", "w21_lo", "w21_di", "w21_up",
                                   "w22_lo", "w22_di", "w22_up")]


class _ExplicitParams(C.Structure):
    _fields_ = [(n, _dp) for n in ("kappa_dry", "kappa_sat_unfrozen", "kappa_sat_frozen", "nu_ss_om",
                                   "nu_ss_quartz", "nu_ss_gravel")] + \
               [(n, C.c_double) for n in ("Omega", "gamma", "gammaT_ref", "alpha", "beta", "T_freeze", "grav")]


class _Aux(C.Structure):
    _fields_ = [(n, _dp) for n in ("theta_l", "kappa", "T", "K", "psi", "Tf_depressed", "total_water",
                                   "total_energy")]


class _CO2Species(C.Structure):
    _fields_ = [("D", _dp), ("theta_eff", _dp), ("c_atm", _dp)]


class _RunoffParams(C.Structure):
    _fields_ = [("f_max", _dp), ("f_over", C.c_double), ("R_sb", C.c_double), ("depth", C.c_double)]


class _Runoff(C.Structure):
    _fields_ = [(n, _dp) for n in ("is_saturated", "h_grad", "infiltration", "R_s", "R_ss", "R_ess")]


# scalar parameters of the explicit stage: EnergyHydrologyParameters defaults pinned by the reference's
# test/standalone/Soil/soil_parameterizations.jl:71-76; T_freeze / grav: ClimaParams 1.1.4 defaults
EXPLICIT_SCALARS = dict(Omega=7.0, gamma=2.64e-2, gammaT_ref=288.0, alpha=0.24, beta=18.3, T_freeze=273.15,
                        grav=9.81)
EXPLICIT_CELL = ("kappa_dry", "kappa_sat_unfrozen", "kappa_sat_frozen", "nu_ss_om", "nu_ss_quartz", "nu_ss_gravel")

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        d = C.c_double
        for name, nargs in [("orc_effective_saturation", 3), ("orc_volumetric_liquid_fraction", 3),
                            ("orc_vg_matric_potential", 4), ("orc_vg_inverse_matric_potential", 4),
                            ("orc_vg_pressure_head", 7), ("orc_vg_dpsidtheta", 7),
                            ("orc_vg_hydraulic_conductivity", 3), ("orc_bc_matric_potential", 3),
                            ("orc_bc_inverse_matric_potential", 3), ("orc_bc_pressure_head", 6),
                            ("orc_bc_dpsidtheta", 6), ("orc_bc_hydraulic_conductivity", 3),
                            ("orc_volumetric_heat_capacity", 7), ("orc_temperature_from_rho_e_int", 6),
                            ("orc_volumetric_internal_energy", 6),
                            ("orc_volumetric_internal_energy_liq", 4), ("orc_heaviside", 2),
                            ("orc_impedance_factor", 2), ("orc_viscosity_factor", 3),
                            ("orc_kappa_sat", 4), ("orc_relative_saturation", 3), ("orc_kersten_number", 7),
                            ("orc_thermal_conductivity", 3), ("orc_thermal_time", 3),
                            ("orc_topmodel_ss_flux", 3), ("orc_topmodel_surface_infiltration", 5)]:
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d] * nargs
        L.orc_soil_Tf_depressed.restype = d
        L.orc_soil_Tf_depressed.argtypes = [C.c_int] + [d] * 12
        L.orc_phase_change_source.restype = d
        L.orc_phase_change_source.argtypes = [C.c_int] + [d] * 14
        pp, ps, pc, pj = (C.POINTER(t) for t in (_Problem, _State, _Cache, _Jacobian))
        px, pa = C.POINTER(_ExplicitParams), C.POINTER(_Aux)
        L.orc_update_aux.argtypes = [pp, px, ps, pa]
        L.orc_update_aux.restype = None
        L.orc_phase_change.argtypes = [pp, px, ps, pa, _dp, _dp]
        L.orc_phase_change.restype = None
        L.orc_update_runoff.argtypes = [pp, px, C.POINTER(_RunoffParams), ps, pa, _dp, C.POINTER(_Runoff)]
        L.orc_update_runoff.restype = None
        L.orc_surface_runoff.argtypes = [pp, px, ps, pa, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_surface_runoff.restype = None
        L.orc_atmos_driven_top_fluxes.argtypes = [pp] + [_dp] * 8
        L.orc_atmos_driven_top_fluxes.restype = None
        L.orc_energy_water_free_drainage.argtypes = [pp, pa, _dp, _dp]
        L.orc_energy_water_free_drainage.restype = None
        pcs = C.POINTER(_CO2Species)
        L.orc_co2_boundary_flux.argtypes = [pp, pcs, _dp, _dp, _dp]
        L.orc_co2_imp_tendency.argtypes = [pp, pcs, _dp, _dp, _dp, _dp]
        L.orc_co2_jacobian.argtypes = [pp, pcs, d, _dp, _dp, _dp, _dp]
        L.orc_co2_implicit_step.argtypes = [pp, pcs, _dp, _dp, _dp, d, C.c_int]
        L.orc_co2_implicit_step.restype = C.c_int
        for n in ("orc_co2_boundary_flux", "orc_co2_imp_tendency", "orc_co2_jacobian"):
            getattr(L, n).restype = None
        L.orc_update_implicit_cache.argtypes = [pp, ps, pc]
        L.orc_update_boundary_fluxes.argtypes = [pp, ps, pc]
        L.orc_compute_imp_tendency.argtypes = [pp, ps, pc, ps]
        L.orc_compute_jacobian.argtypes = [pp, ps, pc, d, pj]
        L.orc_ldiv.argtypes = [pp, pj, ps, ps]
        L.orc_implicit_step.argtypes = [pp, ps, pc, pj, d, C.c_int, d, _dp]
        L.orc_implicit_step.restype = C.c_int
        L.orc_column_integral.argtypes = [pp, _dp, _dp]
        for n in ("orc_update_implicit_cache", "orc_update_boundary_fluxes", "orc_compute_imp_tendency",
                  "orc_compute_jacobian", "orc_ldiv", "orc_column_integral"):
            getattr(L, n).restype = None
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous float64"
    return a.ctypes.data_as(_dp)


CELL_PARAMS = ("nu", "theta_r", "K_sat", "S_s", "hcm_a", "hcm_b", "hcm_m", "rho_c_ds", "K_lag",
               "kappa_lag", "theta_l_lag", "is_saturated")
COL_PARAMS = ("R_ss", "R_ess", "h_grad", "theta_bc_top", "theta_bc_bot")


class Problem:
    """Holds the inputs of the implicit path; scalars are broadcast to fields."""

    def __init__(self, *, model, closure=VAN_GENUCHTEN, top_bc=TOP_FLUX, bottom_bc=BOT_FLUX,
                 has_topmodel_source=False, z_f, ncol, z_c=None, nthreads=1, earth=None, **fields):
        self.model, self.closure, self.top_bc, self.bottom_bc = model, closure, top_bc, bottom_bc
        self.has_topmodel_source = bool(has_topmodel_source)
        self.z_f = np.ascontiguousarray(z_f, dtype=np.float64)
        self.N = self.z_f.size - 1
        self.z_c = (np.ascontiguousarray(z_c, dtype=np.float64) if z_c is not None
                    else 0.5 * (self.z_f[1:] + self.z_f[:-1]))
        self.ncol = int(ncol)
        self.nthreads = int(nthreads)
        self.earth = dict(EARTH if earth is None else earth)
        self.f = {}
        for k, v in fields.items():
            self.set(k, v)

    def set(self, name, v):
        if name in CELL_PARAMS:
            shape = (self.ncol, self.N)
        elif name in COL_PARAMS:
            shape = (self.ncol,)
        else:
            raise KeyError(name)
        a = np.empty(shape, dtype=np.float64)
        a[...] = v
        self.f[name] = a

    def c_struct(self):
        P = _Problem()
        P.model, P.closure, P.top_bc, P.bottom_bc = self.model, self.closure, self.top_bc, self.bottom_bc
        P.has_topmodel_source = int(self.has_topmodel_source)
        P.N, P.ncol, P.nthreads = self.N, self.ncol, self.nthreads
        P.z_c, P.z_f = _ptr(self.z_c), _ptr(self.z_f)
        for k in CELL_PARAMS + COL_PARAMS:
            setattr(P, k, _ptr(self.f.get(k)))
        for k, v in self.earth.items():
            setattr(P, k, v)
        return P

    # ---- allocation helpers -------------------------------------------------
    def new_state(self):
        return State(self)

    def new_cache(self):
        return Cache(self)

    def new_jacobian(self):
        return Jacobian(self)

    # ---- hooks ---------------------------------------------------------------
    def update_implicit_cache(self, Y, p):
        P, y, c = self.c_struct(), Y.c_struct(), p.c_struct()
        lib().orc_update_implicit_cache(C.byref(P), C.byref(y), C.byref(c))

    def update_boundary_fluxes(self, Y, p):
        P, y, c = self.c_struct(), Y.c_struct(), p.c_struct()
        lib().orc_update_boundary_fluxes(C.byref(P), C.byref(y), C.byref(c))

    def compute_imp_tendency(self, dY, Y, p):
        P, y, c, d = self.c_struct(), Y.c_struct(), p.c_struct(), dY.c_struct()
        lib().orc_compute_imp_tendency(C.byref(P), C.byref(y), C.byref(c), C.byref(d))

    def compute_jacobian(self, W, Y, p, dtgamma):
        P, y, c, w = self.c_struct(), Y.c_struct(), p.c_struct(), W.c_struct()
        lib().orc_compute_jacobian(C.byref(P), C.byref(y), C.byref(c), float(dtgamma), C.byref(w))

    def ldiv(self, x, W, b):
        P, w, bb, xx = self.c_struct(), W.c_struct(), b.c_struct(), x.c_struct()
        lib().orc_ldiv(C.byref(P), C.byref(w), C.byref(bb), C.byref(xx))

    def implicit_step(self, U, dtgamma, max_iters, tol=-1.0, p=None, W=None):
        """Advance U in place through one implicit ARS111 stage; returns (iters, ||dx||)."""
        p = p or self.new_cache()
        W = W or self.new_jacobian()
        P, u, c, w = self.c_struct(), U.c_struct(), p.c_struct(), W.c_struct()
        nrm = C.c_double(0.0)
        it = lib().orc_implicit_step(C.byref(P), C.byref(u), C.byref(c), C.byref(w), float(dtgamma),
                                     int(max_iters), float(tol), C.byref(nrm))
        return it, nrm.value

    # ---- explicit stage of EnergyHydrology -------------------------------
    def explicit_params(self, scalars=None, **cell_fields):
        """EnergyHydrologyParameters the explicit stage reads: six per-cell fields + scalars."""
        return ExplicitParams(self, scalars, **cell_fields)

    def new_aux(self):
        return Aux(self)

    def update_aux(self, X, Y, a):
        P, x, y, aa = self.c_struct(), X.c_struct(), Y.c_struct(), a.c_struct()
        lib().orc_update_aux(C.byref(P), C.byref(x), C.byref(y), C.byref(aa))

    def phase_change(self, X, Y, a, dtheta_l, dtheta_i):
        """source!(dY, ::PhaseChange, ...): ADDS into dtheta_l, dtheta_i (C-contiguous (ncol, N))."""
        P, x, y, aa = self.c_struct(), X.c_struct(), Y.c_struct(), a.c_struct()
        lib().orc_phase_change(C.byref(P), C.byref(x), C.byref(y), C.byref(aa), _ptr(dtheta_l), _ptr(dtheta_i))

    def update_runoff(self, Y, precip, f_max, f_over, R_sb, depth, X=None, a=None):
        """update_infiltration_water_flux!(p, ::TOPMODELRunoff, input, Y, t, model) -> Runoff bundle."""
        out = Runoff(self)
        fm = np.ascontiguousarray(np.broadcast_to(np.asarray(f_max, dtype=np.float64), (self.ncol,)))
        pr = np.ascontiguousarray(np.broadcast_to(np.asarray(precip, dtype=np.float64), (self.ncol,)))
        R = _RunoffParams(_ptr(fm), float(f_over), float(R_sb), float(depth))
        P, y, o = self.c_struct(), Y.c_struct(), out.c_struct()
        x = X.c_struct() if X is not None else None
        aa = a.c_struct() if a is not None else None
        lib().orc_update_runoff(C.byref(P), C.byref(x) if x is not None else None, C.byref(R), C.byref(y),
                                C.byref(aa) if aa is not None else None, _ptr(pr), C.byref(o))
        return out

    # ---- atmosphere-driven top boundary fluxes ------------------------------
    def surface_runoff(self, Y, kind, liquid_influx, X=None, a=None):
        """update_infiltration_water_flux! of NoRunoff (kind 0) / SurfaceRunoff (kind 1) -> (is_saturated, infiltration, R_s)"""
        inp = np.ascontiguousarray(np.broadcast_to(np.asarray(liquid_influx, dtype=np.float64), (self.ncol,)))
        sat, inf, R_s = np.zeros((self.ncol, self.N)), np.zeros(self.ncol), np.zeros(self.ncol)
        P, y = self.c_struct(), Y.c_struct()
        x = X.c_struct() if X is not None else None
        aa = a.c_struct() if a is not None else None
        lib().orc_surface_runoff(C.byref(P), C.byref(x) if x is not None else None, C.byref(y),
                                 C.byref(aa) if aa is not None else None, int(kind), _ptr(inp), _ptr(sat), _ptr(inf), _ptr(R_s))
        return sat, inf, R_s

    def atmos_driven_top_fluxes(self, infiltration, vapor_flux_liq, lhf, shf, R_n, T_air):
        """soil_boundary_fluxes!(::AtmosDrivenFluxBC, Val((:soil,)), ...) after the runoff -> (top_bc.water, top_bc.heat)"""
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.ncol,)))
                for v in (infiltration, vapor_flux_liq, lhf, shf, R_n, T_air)]
        w, h = np.zeros(self.ncol), np.zeros(self.ncol)
        P = self.c_struct()
        lib().orc_atmos_driven_top_fluxes(C.byref(P), *[_ptr(v) for v in arrs], _ptr(w), _ptr(h))
        return w, h

    def energy_water_free_drainage(self, a):
        """soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...) -> (bottom_bc.water, bottom_bc.heat)"""
        w, h = np.zeros(self.ncol), np.zeros(self.ncol)
        P, aa = self.c_struct(), a.c_struct()
        lib().orc_energy_water_free_drainage(C.byref(P), C.byref(aa), _ptr(w), _ptr(h))
        return w, h

    # ---- SoilCO2Model implicit diffusion -----------------------------------
    def co2_species(self, D, theta_eff, c_atm=None):
        return CO2Species(self, D, theta_eff, c_atm)

    def co2_boundary_flux(self, S, Cc, top_bc, dflux):
        P, s = self.c_struct(), S.c_struct()
        lib().orc_co2_boundary_flux(C.byref(P), C.byref(s), _ptr(Cc), _ptr(top_bc), _ptr(dflux))

    def co2_imp_tendency(self, S, Cc, top_bc, bot_bc):
        out = np.zeros_like(Cc)
        P, s = self.c_struct(), S.c_struct()
        lib().orc_co2_imp_tendency(C.byref(P), C.byref(s), _ptr(Cc), _ptr(top_bc), _ptr(bot_bc), _ptr(out))
        return out

    def co2_jacobian(self, S, dtgamma, dflux=None):
        lo, di, up = (np.zeros((self.ncol, self.N)) for _ in range(3))
        P, s = self.c_struct(), S.c_struct()
        lib().orc_co2_jacobian(C.byref(P), C.byref(s), float(dtgamma), _ptr(dflux), _ptr(lo), _ptr(di), _ptr(up))
        return lo, di, up

    def co2_implicit_step(self, S, Cc, top_bc, bot_bc, dtgamma, max_iters):
        """advances Cc (and, for a state BC, top_bc) in place"""
        P, s = self.c_struct(), S.c_struct()
        return lib().orc_co2_implicit_step(C.byref(P), C.byref(s), _ptr(Cc), _ptr(top_bc), _ptr(bot_bc),
                                           float(dtgamma), int(max_iters))

    def column_integral(self, field):
        out = np.zeros(self.ncol)
        P = self.c_struct()
        lib().orc_column_integral(C.byref(P), _ptr(np.ascontiguousarray(field)), _ptr(out))
        return out


class _Bundle:
    cell, col, ctype = (), (), None

    def __init__(self, prob):
        for k in self.cell:
            setattr(self, k, np.zeros((prob.ncol, prob.N)))
        for k in self.col:
            setattr(self, k, np.zeros(prob.ncol))

    def c_struct(self):
        s = self.ctype()
        for k in self.cell + self.col:
            setattr(s, k, _ptr(getattr(self, k)))
        return s

    def copy(self):
        import copy
        new = copy.copy(self)
        for k in self.cell + self.col:
            setattr(new, k, getattr(self, k).copy())
        return new


class State(_Bundle):
    cell, col, ctype = ("theta_l", "rho_e_int", "theta_i"), ("intF_w", "intF_e"), _State


class Cache(_Bundle):
    cell = ("K", "psi", "T")
    col = ("top_bc_w", "bot_bc_w", "top_bc_h", "bot_bc_h", "dfluxBCdY", "total_water")
    ctype = _Cache


class Aux(_Bundle):
    cell = ("theta_l", "kappa", "T", "K", "psi", "Tf_depressed")
    col, ctype = ("total_water", "total_energy"), _Aux


class CO2Species:
    def __init__(self, prob, D, theta_eff, c_atm=None):
        self.D, self.theta_eff = np.empty((prob.ncol, prob.N)), np.empty((prob.ncol, prob.N))
        self.D[...] = D
        self.theta_eff[...] = theta_eff
        self.c_atm = None
        if c_atm is not None:
            self.c_atm = np.empty(prob.ncol)
            self.c_atm[...] = c_atm

    def c_struct(self):
        return _CO2Species(_ptr(self.D), _ptr(self.theta_eff), _ptr(self.c_atm))


class Runoff(_Bundle):
    cell, col, ctype = ("is_saturated",), ("h_grad", "infiltration", "R_s", "R_ss", "R_ess"), _Runoff


class ExplicitParams:
    def __init__(self, prob, scalars=None, **cell_fields):
        self.scalars = dict(EXPLICIT_SCALARS)
        self.scalars.update(scalars or {})
        self.f = {}
        for k in EXPLICIT_CELL:
            a = np.empty((prob.ncol, prob.N))
            a[...] = cell_fields[k]
            self.f[k] = a

    def c_struct(self):
        s = _ExplicitParams()
        for k, v in self.f.items():
            setattr(s, k, _ptr(v))
        for k, v in self.scalars.items():
            setattr(s, k, float(v))
        return s


class Jacobian(_Bundle):
    cell = tuple(f"w{b}_{d}" for b in ("11", "21", "22") for d in ("lo", "di", "up"))
    col, ctype = (), _Jacobian
