"""`SoilColumnSolver`: object wrapper over one `clb_handle` (one GPU, one shard of
columns).  Arrays cross this boundary in the reference's layout -- per-cell fields
are `(ncol, N)` C-contiguous float64 (level fastest, level 0 = bottom), per-column
fields `(ncol,)` -- as numpy arrays (host) or CUDA torch tensors (device); the
library transposes into its column-fastest mirrors on the device.  All arithmetic
happens in libclimaland_b200.so; nothing here computes."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import K, ClbError, Config, ExplicitParams, RunoffParams, Stats, check

# LandParameters constants (ClimaParams defaults; passed in as numbers,
# src/shared_utilities/Parameters.jl:86-105)
EARTH = dict(rho_l=1000.0, rho_i=916.7, cp_l=4181.0, cp_i=2100.0, T_ref=273.16,
             LH_f0=2.8344e6 - 2.5008e6)

RICHARDS, ENERGY_HYDROLOGY = K["CLB_RICHARDS"], K["CLB_ENERGY_HYDROLOGY"]
VAN_GENUCHTEN, BROOKS_COREY = K["CLB_VAN_GENUCHTEN"], K["CLB_BROOKS_COREY"]
TOP_FLUX, TOP_MOISTURE_STATE = K["CLB_TOP_FLUX"], K["CLB_TOP_MOISTURE_STATE"]
BOT_FLUX, BOT_FREE_DRAINAGE, BOT_MOISTURE_STATE = (K["CLB_BOT_FLUX"], K["CLB_BOT_FREE_DRAINAGE"],
                                                   K["CLB_BOT_MOISTURE_STATE"])
MATH_FAST, MATH_LIBM = K["CLB_MATH_FAST"], K["CLB_MATH_LIBM"]
VARIANT_AUTO, VARIANT_REGISTER_COLUMN, VARIANT_GENERIC, VARIANT_LANE_PER_CELL, VARIANT_LANE_QUAD, VARIANT_LANE_QUAD_PIPELINED = (
    K["CLB_VARIANT_AUTO"], K["CLB_VARIANT_REGISTER_COLUMN"], K["CLB_VARIANT_GENERIC"], K["CLB_VARIANT_LANE_PER_CELL"],
    K["CLB_VARIANT_LANE_QUAD"], K["CLB_VARIANT_LANE_QUAD_PIPELINED"])
VARIANT_LANE_OCTET = K["CLB_VARIANT_LANE_OCTET"]
LAYOUT_AUTO, LAYOUT_COLUMN_FASTEST, LAYOUT_LEVEL_FASTEST = (K["CLB_LAYOUT_AUTO"], K["CLB_LAYOUT_COLUMN_FASTEST"],
                                                            K["CLB_LAYOUT_LEVEL_FASTEST"])

FIELDS = {k[len("CLB_F_"):].lower(): v for k, v in K.items()
          if k.startswith("CLB_F_") and k not in ("CLB_F_NUM", "CLB_F_NUM_CELL")}
NUM_CELL = K["CLB_F_NUM_CELL"]


def field_id(name):
    try:
        return FIELDS[name.lower()]
    except KeyError:
        raise KeyError(f"unknown field {name!r}; known: {sorted(FIELDS)}") from None


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class SoilColumnSolver:
    def __init__(self, *, model, n_columns, z_f, z_c=None, closure=VAN_GENUCHTEN, top_bc=TOP_FLUX,
                 bottom_bc=BOT_FLUX, has_topmodel_source=False, device=0, stream=None, math_mode=MATH_FAST,
                 kernel_variant=VARIANT_AUTO, layout=LAYOUT_AUTO, earth=None, active_columns=None,
                 n_columns_total=None, out_of_place=False):
        self.L = _lib.lib()
        z_f = np.ascontiguousarray(z_f, dtype=np.float64)
        self.N = int(z_f.size - 1)
        z_c = (np.ascontiguousarray(z_c, dtype=np.float64) if z_c is not None else 0.5 * (z_f[1:] + z_f[:-1]))
        self.z_f, self.z_c = z_f, z_c
        self.model, self.closure = model, closure
        self.ncol = int(n_columns)
        self.ncol_total = int(n_columns_total if n_columns_total is not None else n_columns)
        e = dict(EARTH if earth is None else earth)
        cfg = Config(abi_version=K["CLB_ABI_VERSION"], model=model, closure=closure, top_bc=top_bc,
                     bottom_bc=bottom_bc, has_topmodel_source=int(bool(has_topmodel_source)), n_levels=self.N,
                     device=int(device), n_columns=self.ncol, stream=stream, math_mode=math_mode,
                     kernel_variant=kernel_variant, layout=layout, **e)
        self.cfg = cfg
        self.h = C.c_void_p()
        check(self.L.clb_create(C.byref(self.h), C.byref(cfg)))
        check(self.L.clb_set_grid(self.h, z_c.ctypes.data_as(_lib._dp), z_f.ctypes.data_as(_lib._dp)))
        if out_of_place:
            check(self.L.clb_set_option(self.h, K["CLB_OPT_OUT_OF_PLACE"], 1))
        if active_columns is not None:
            idx = np.ascontiguousarray(active_columns, dtype=np.int64)
            check(self.L.clb_set_active_columns(self.h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size))

    @classmethod
    def from_workload(cls, w, closure=VAN_GENUCHTEN, top_bc=TOP_FLUX, bottom_bc=BOT_FLUX, **kw):
        """A solver with every field of a `workloads.make_workload` dict uploaded (bench, smoke, tests)."""
        model = RICHARDS if w["model"] == "richards" else ENERGY_HYDROLOGY
        s = cls(model=model, n_columns=w["ncol"], z_f=w["z_f"], z_c=w["z_c"], closure=closure, top_bc=top_bc,
                bottom_bc=bottom_bc, has_topmodel_source=w.get("topmodel", False), **kw)
        for k, v in w.items():
            if k.lower() in FIELDS:
                s.set(k, v)
        return s

    def set_option(self, name, value):
        """clb_set_option by name: "host_route", "host_chunks", "tile_boxes", "out_of_place", ..."""
        check(self.L.clb_set_option(self.h, K["CLB_OPT_" + name.upper()], int(value)))

    # ---- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.clb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.clb_sync(self.h))

    # ---- field transfer ----------------------------------------------------
    def _shape(self, fid):
        return (self.ncol_total, self.N) if fid < NUM_CELL else (self.ncol_total,)

    def set(self, name, value):
        """Upload a field (array in the reference layout) or broadcast a scalar."""
        fid = field_id(name)
        if np.isscalar(value):
            check(self.L.clb_fill_field(self.h, fid, float(value)))
            return
        if _is_torch(value):
            import torch
            assert value.is_cuda and value.dtype == torch.float64 and value.is_contiguous()
            assert tuple(value.shape) == self._shape(fid), (name, tuple(value.shape), self._shape(fid))
            ptr, mem = C.c_void_p(value.data_ptr()), K["CLB_DEVICE"]
        else:
            a = np.ascontiguousarray(value, dtype=np.float64)
            if a.shape != self._shape(fid):
                a = np.ascontiguousarray(np.broadcast_to(a, self._shape(fid)))
            ptr, mem = C.c_void_p(a.ctypes.data), K["CLB_HOST"]
        sc = self.N if fid < NUM_CELL else 1
        check(self.L.clb_set_field(self.h, fid, ptr, 1, sc, mem))
        if mem == K["CLB_HOST"]:
            self.sync()  # the temporary above must outlive the copy

    def get(self, name, out=None):
        """Download a field into `out` (numpy or CUDA torch tensor, reference layout)."""
        fid = field_id(name)
        if out is None:
            out = np.zeros(self._shape(fid))
        if _is_torch(out):
            ptr, mem = C.c_void_p(out.data_ptr()), K["CLB_DEVICE"]
        else:
            assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.shape == self._shape(fid)
            ptr, mem = C.c_void_p(out.ctypes.data), K["CLB_HOST"]
        sc = self.N if fid < NUM_CELL else 1
        check(self.L.clb_get_field(self.h, fid, ptr, 1, sc, mem))
        return out

    def axpy(self, y, a, x):
        """y <- y + a x on the device mirrors (the integrator's explicit update)."""
        check(self.L.clb_field_axpy(self.h, field_id(y), float(a), field_id(x)))

    def ldiv_diagonal(self, w, b, x):
        """x <- b / w entry by entry: a DiagonalMatrixRow block (surface variables of an integrated model)."""
        check(self.L.clb_ldiv_diagonal(self.h, field_id(w), field_id(b), field_id(x)))

    def copy(self, dst, src):
        check(self.L.clb_field_copy(self.h, field_id(dst), field_id(src)))

    def device_ptr(self, name):
        p, sl, sc = C.c_void_p(), C.c_int64(), C.c_int64()
        check(self.L.clb_field_device_ptr(self.h, field_id(name), C.byref(p), C.byref(sl), C.byref(sc)))
        return p.value, sl.value, sc.value

    # ---- hooks ---------------------------------------------------------------
    def update_implicit_cache(self):
        check(self.L.clb_update_implicit_cache(self.h))

    def update_boundary_fluxes(self):
        check(self.L.clb_update_boundary_fluxes(self.h))

    def compute_imp_tendency(self):
        check(self.L.clb_compute_imp_tendency(self.h))

    def compute_jacobian(self, dtgamma):
        check(self.L.clb_compute_jacobian(self.h, float(dtgamma)))

    def ldiv(self):
        check(self.L.clb_ldiv(self.h))

    # ---- explicit stage of EnergyHydrology -------------------
    def set_explicit_params(self, *, Omega, gamma, gammaT_ref, alpha, beta, T_freeze, grav):
        """Scalars of EnergyHydrologyParameters (energy_hydrology.jl:150-160) + T_freeze, grav."""
        p = ExplicitParams(Omega, gamma, gammaT_ref, alpha, beta, T_freeze, grav)
        check(self.L.clb_set_explicit_params(self.h, C.byref(p)))

    # ---- SoilCO2Model implicit diffusion ---------------------
    def set_co2_top_state(self, co2=False, o2=False):
        """AtmosCO2StateBC / AtmosO2StateBC at the top (values co2_c_atm / o2_c_atm) instead of flux values."""
        check(self.L.clb_set_option(self.h, K["CLB_OPT_CO2_TOP_STATE"], int(co2)))
        check(self.L.clb_set_option(self.h, K["CLB_OPT_O2_TOP_STATE"], int(o2)))

    def soilco2_update_boundary_fluxes(self):
        check(self.L.clb_soilco2_update_boundary_fluxes(self.h))

    def soilco2_compute_imp_tendency(self):
        check(self.L.clb_soilco2_compute_imp_tendency(self.h))

    def soilco2_compute_jacobian(self, dtgamma):
        check(self.L.clb_soilco2_compute_jacobian(self.h, float(dtgamma)))

    def soilco2_implicit_step(self, dtgamma, max_iters):
        check(self.L.clb_soilco2_implicit_step(self.h, float(dtgamma), int(max_iters)))

    def set_runoff_params(self, *, f_over, R_sb, depth):
        """TOPMODELRunoff scalars (Runoff/Runoff.jl:190-222) and the domain depth."""
        p = RunoffParams(f_over, R_sb, depth)
        check(self.L.clb_set_runoff_params(self.h, C.byref(p)))

    def update_runoff(self):
        check(self.L.clb_update_runoff(self.h))

    def update_aux(self):
        check(self.L.clb_update_aux(self.h))

    def phase_change_source(self):
        check(self.L.clb_phase_change_source(self.h))

    def update_aux_and_phase_change(self):
        check(self.L.clb_update_aux_and_phase_change(self.h))

    def implicit_step(self, dtgamma, max_iters, tol=-1.0, want_stats=False):
        st = Stats() if want_stats else None
        check(self.L.clb_implicit_step(self.h, float(dtgamma), int(max_iters), float(tol),
                                       C.byref(st) if st is not None else None))
        if st is not None:
            return dict(iterations=st.iterations, converged=bool(st.converged), dx_norm=st.dx_norm,
                        nan_count=st.nan_count)
        return None

    def last_variant(self):
        """CLB_VARIANT_* the last implicit_step launched (what VARIANT_AUTO resolved to)."""
        v = C.c_int32(0)
        check(self.L.clb_last_variant(self.h, C.byref(v)))
        return int(v.value)

    def implicit_step_host(self, dtgamma, max_iters, inputs, outputs):
        """inputs / outputs: dict name -> contiguous float64 numpy array (reference layout,
        ideally pinned).  One call = upload, fused stage, download, sync."""
        n_in, n_out = len(inputs), len(outputs)
        fi = (C.c_int32 * n_in)(*[field_id(k) for k in inputs])
        pi = (C.c_void_p * n_in)(*[v.ctypes.data for v in inputs.values()])
        fo = (C.c_int32 * n_out)(*[field_id(k) for k in outputs])
        po = (C.c_void_p * n_out)(*[v.ctypes.data for v in outputs.values()])
        check(self.L.clb_implicit_step_host(self.h, float(dtgamma), int(max_iters), fi, pi, n_in, fo, po, n_out))

    def soil_step_host(self, dt, max_iters, inputs, outputs):
        """A whole EnergyHydrology soil step from / to host arrays holding the state at t_n (clb_soil_step_host):
        update_aux! + PhaseChange, TOPMODEL runoff, u + dt T_exp(u) and the implicit stage on the device; only the
        fields in `inputs` / `outputs` cross PCIe (dict name -> contiguous float64 numpy array, ideally pinned)."""
        n_in, n_out = len(inputs), len(outputs)
        fi = (C.c_int32 * n_in)(*[field_id(k) for k in inputs])
        pi = (C.c_void_p * n_in)(*[v.ctypes.data for v in inputs.values()])
        fo = (C.c_int32 * n_out)(*[field_id(k) for k in outputs])
        po = (C.c_void_p * n_out)(*[v.ctypes.data for v in outputs.values()])
        check(self.L.clb_soil_step_host(self.h, float(dt), int(max_iters), fi, pi, n_in, fo, po, n_out))

    def update_atmos_driven_fluxes(self, runoff_model=2):
        """soil_boundary_fluxes!(::AtmosDrivenFluxBC, ...) after the host's turbulent fluxes and net radiation:
        runoff (0 NoRunoff, 1 SurfaceRunoff, 2 TOPMODELRunoff) + the assembly of top_bc (clb_update_atmos_driven_fluxes)"""
        check(self.L.clb_update_atmos_driven_fluxes(self.h, int(runoff_model)))

    def update_energy_water_free_drainage(self):
        """soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...)"""
        check(self.L.clb_update_energy_water_free_drainage(self.h))

    def ldiv_all(self, soil=True, soilco2=False, surface=False):
        """ldiv! of an integrated model's FieldMatrixWithSolver in one launch (clb_ldiv_all): the soil blocks, the SoilCO2
        tridiagonals and the DiagonalMatrixRow block of one surface variable, whichever are present"""
        check(self.L.clb_ldiv_all(self.h, (1 if soil else 0) | (2 if soilco2 else 0) | (4 if surface else 0)))

    def soil_step(self, dt, max_iters=3):
        """A whole EnergyHydrology soil step on resident state (clb_soil_step): explicit cells, the per-column sweep
        (runoff, column integrals, explicit update), the fused implicit stage -- three launches, no host transfer."""
        check(self.L.clb_soil_step(self.h, float(dt), int(max_iters)))

    def column_integral(self, cell_field, col_field_out):
        check(self.L.clb_column_integral(self.h, field_id(cell_field), field_id(col_field_out)))

    def global_balance(self):
        out = np.zeros(4)
        check(self.L.clb_global_balance(self.h, out.ctypes.data_as(_lib._dp)))
        return out

    # ---- multi-GPU -----------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        check(_lib.lib().clb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, n_ranks, rank):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(self.L.clb_comm_init(self.h, buf, int(n_ranks), int(rank)))
