"""ctypes binding of libclimaland_b200.so -- the same C ABI ClimaLand's Julia hooks
`ccall` (include/climaland_b200.h).  There is no fallback: if the library cannot
be built or loaded, importing this module's `lib()` raises."""
import ctypes as C
import os
import re

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "climaland_b200.h")

_dp = C.POINTER(C.c_double)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("closure", C.c_int32),
        ("top_bc", C.c_int32), ("bottom_bc", C.c_int32), ("has_topmodel_source", C.c_int32),
        ("n_levels", C.c_int32), ("device", C.c_int32), ("n_columns", C.c_int64),
        ("stream", C.c_void_p), ("math_mode", C.c_int32), ("kernel_variant", C.c_int32),
        ("rho_l", C.c_double), ("rho_i", C.c_double), ("cp_l", C.c_double),
        ("cp_i", C.c_double), ("T_ref", C.c_double), ("LH_f0", C.c_double),
        ("layout", C.c_int32), ("reserved", C.c_int32),
    ]


class ExplicitParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Omega", "gamma", "gammaT_ref", "alpha", "beta", "T_freeze", "grav")]


class RunoffParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("f_over", "R_sb", "depth")]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("dx_norm", C.c_double),
                ("nan_count", C.c_int64)]


def _parse_enums():
    """Field ids and enum constants come from the header, so Python can never drift from the ABI."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    consts = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = (x.strip() for x in item.split("=", 1))
                nxt = consts[val] if val in consts else int(val, 0)
            else:
                name = item
            consts[name] = nxt
            nxt += 1
    m = re.search(r"#define\s+CLB_ABI_VERSION\s+(\d+)", src)
    consts["CLB_ABI_VERSION"] = int(m.group(1))
    return consts


K = _parse_enums()
EXPORTS = re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S))
EXPORTS = sorted(set(EXPORTS))

_lib = None


class ClbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"climaland_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    # CLB_LIBRARY_PATH: load an alternative build of the same ABI (kernel tuning experiments)
    path = os.environ.get("CLB_LIBRARY_PATH") or _build.build()
    L = C.CDLL(path)
    h = C.c_void_p
    i32, i64, d = C.c_int32, C.c_int64, C.c_double
    L.clb_abi_version.restype = C.c_int
    L.clb_last_error.restype = C.c_char_p
    sig = {
        "clb_create": [C.POINTER(h), C.POINTER(Config)],
        "clb_destroy": [h], "clb_sync": [h], "clb_set_stream": [h, C.c_void_p],
        "clb_set_grid": [h, _dp, _dp],
        "clb_set_active_columns": [h, C.POINTER(i64), i64],
        "clb_set_field": [h, i32, C.c_void_p, i64, i64, i32],
        "clb_get_field": [h, i32, C.c_void_p, i64, i64, i32],
        "clb_fill_field": [h, i32, d],
        "clb_field_axpy": [h, i32, d, i32], "clb_field_copy": [h, i32, i32], "clb_ldiv_diagonal": [h, i32, i32, i32],
        "clb_field_device_ptr": [h, i32, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(i64)],
        "clb_set_option": [h, i32, i64],
        "clb_update_implicit_cache": [h], "clb_update_boundary_fluxes": [h],
        "clb_compute_imp_tendency": [h], "clb_compute_jacobian": [h, d], "clb_ldiv": [h],
        "clb_set_explicit_params": [h, C.POINTER(ExplicitParams)],
        "clb_set_runoff_params": [h, C.POINTER(RunoffParams)], "clb_update_runoff": [h],
        "clb_soilco2_update_boundary_fluxes": [h], "clb_soilco2_compute_imp_tendency": [h],
        "clb_soilco2_compute_jacobian": [h, d], "clb_soilco2_implicit_step": [h, d, i32],
        "clb_update_aux": [h], "clb_phase_change_source": [h], "clb_update_aux_and_phase_change": [h],
        "clb_implicit_step": [h, d, i32, d, C.POINTER(Stats)],
        "clb_implicit_step_host": [h, d, i32, C.POINTER(i32), C.POINTER(C.c_void_p), i32,
                                   C.POINTER(i32), C.POINTER(C.c_void_p), i32],
        "clb_soil_step_host": [h, d, i32, C.POINTER(i32), C.POINTER(C.c_void_p), i32,
                               C.POINTER(i32), C.POINTER(C.c_void_p), i32],
        "clb_soil_step": [h, d, i32],
        "clb_ldiv_all": [h, C.c_uint32],
        "clb_update_atmos_driven_fluxes": [h, i32],
        "clb_update_energy_water_free_drainage": [h],
        "clb_column_integral": [h, i32, i32],
        "clb_global_balance": [h, _dp],
        "clb_test_math": [i32, _dp, _dp, _dp, i64],
        "clb_last_variant": [h, C.POINTER(C.c_int32)],
        "clb_comm_unique_id": [C.c_void_p],
        "clb_comm_init": [h, C.c_void_p, i32, i32],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    if L.clb_abi_version() != K["CLB_ABI_VERSION"]:
        raise RuntimeError("libclimaland_b200.so ABI version does not match include/climaland_b200.h")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise ClbError(rc, lib().clb_last_error().decode())
