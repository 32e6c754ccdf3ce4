"""Error behaviour of the C ABI on a GPU box (SURVEY 8b "Errors"): every misuse is a negative clb_status with a
message from clb_last_error(), never a crash or an exception across the boundary; the Python / Julia bindings turn
the status into their own exception, which is how the reference's hooks report errors."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cl():
    import climaland_b200 as cl
    return cl


def _cfg(cl, **kw):
    K = cl._lib.K
    d = dict(abi_version=K["CLB_ABI_VERSION"], model=K["CLB_RICHARDS"], closure=0, top_bc=0, bottom_bc=0,
             has_topmodel_source=0, n_levels=15, device=0, n_columns=8, stream=None, math_mode=0, kernel_variant=0,
             layout=0, **cl.EARTH)
    d.update(kw)
    return cl._lib.Config(**d)


@pytest.mark.parametrize("kw,code", [
    (dict(abi_version=1), "CLB_ERR_INVALID"), (dict(model=7), "CLB_ERR_INVALID"), (dict(closure=3), "CLB_ERR_INVALID"),
    (dict(n_columns=0), "CLB_ERR_INVALID"), (dict(n_levels=1), "CLB_ERR_INVALID"), (dict(n_levels=100000), "CLB_ERR_INVALID"),
    (dict(device=99), "CLB_ERR_INVALID"), (dict(kernel_variant=42), "CLB_ERR_INVALID"), (dict(layout=9), "CLB_ERR_INVALID"),
    (dict(top_bc=5), "CLB_ERR_INVALID"), (dict(math_mode=2), "CLB_ERR_INVALID")])
def test_create_rejects_bad_configurations(kw, code):
    cl = _cl()
    L = cl._lib.lib()
    h = C.c_void_p()
    rc = L.clb_create(C.byref(h), C.byref(_cfg(cl, **kw)))
    assert rc == cl._lib.K[code] and not h.value
    assert len(L.clb_last_error()) > 10


def test_null_and_unset_arguments_are_statuses():
    cl = _cl()
    L, K = cl._lib.lib(), cl._lib.K
    assert L.clb_create(None, None) == K["CLB_ERR_INVALID"]
    assert L.clb_sync(None) == K["CLB_ERR_INVALID"]
    assert L.clb_destroy(None) == K["CLB_OK"]                       # destroying nothing is fine
    h = C.c_void_p()
    assert L.clb_create(C.byref(h), C.byref(_cfg(cl))) == K["CLB_OK"]
    assert L.clb_set_grid(h, None, None) == K["CLB_ERR_INVALID"]
    assert L.clb_set_field(h, K["CLB_F_NU"], None, 1, 15, K["CLB_HOST"]) == K["CLB_ERR_INVALID"]
    assert L.clb_set_field(h, 9999, C.c_void_p(8), 1, 15, K["CLB_HOST"]) == K["CLB_ERR_INVALID"]
    a = np.zeros((8, 15))
    assert L.clb_set_field(h, K["CLB_F_NU"], a.ctypes.data, -1, 15, K["CLB_HOST"]) == K["CLB_ERR_INVALID"]
    assert L.clb_set_field(h, K["CLB_F_NU"], a.ctypes.data, 1, 15, 7) == K["CLB_ERR_INVALID"]
    assert L.clb_get_field(h, K["CLB_F_P_K"], a.ctypes.data, 1, 15, K["CLB_HOST"]) == K["CLB_ERR_UNSET"]
    # hooks before the grid / the fields exist
    assert L.clb_update_implicit_cache(h) == K["CLB_ERR_UNSET"]
    assert L.clb_implicit_step(h, 900.0, 3, -1.0, None) == K["CLB_ERR_UNSET"]
    assert L.clb_implicit_step(h, 900.0, 0, -1.0, None) == K["CLB_ERR_INVALID"]
    assert L.clb_set_option(h, 12345, 1) == K["CLB_ERR_INVALID"]
    assert L.clb_set_explicit_params(h, None) == K["CLB_ERR_INVALID"]
    assert L.clb_set_runoff_params(h, None) == K["CLB_ERR_INVALID"]
    assert L.clb_update_runoff(h) == K["CLB_ERR_UNSET"]
    assert L.clb_soilco2_implicit_step(h, 900.0, 0) == K["CLB_ERR_INVALID"]
    assert L.clb_global_balance(h, None) == K["CLB_ERR_INVALID"]
    assert L.clb_last_variant(h, None) == K["CLB_ERR_INVALID"]
    assert L.clb_comm_unique_id(None) == K["CLB_ERR_INVALID"]
    assert b"clb_comm_unique_id" in L.clb_last_error()
    assert L.clb_destroy(h) == K["CLB_OK"]


def test_variant_requests_that_do_not_apply_are_refused():
    cl = _cl()
    from climaland_b200 import workloads
    from helpers import cuda_solver
    w = workloads.make_workload("richards", 32, N=20, seed=1)
    for kv in (cl.VARIANT_REGISTER_COLUMN, cl.VARIANT_LANE_QUAD, cl.VARIANT_LANE_QUAD_PIPELINED):
        s = cuda_solver(w, kernel_variant=kv)
        with pytest.raises(cl.ClbError) as e:
            s.implicit_step(1800.0, 2)
        assert e.value.code == cl._lib.K["CLB_ERR_INVALID"]
        s.close()
    # the lane octet covers 15 <= N <= 64
    for N in (10, 14):
        s = cuda_solver(workloads.make_workload("richards", 32, N=N, seed=1), kernel_variant=cl.VARIANT_LANE_OCTET)
        with pytest.raises(cl.ClbError) as e:
            s.implicit_step(1800.0, 2)
        assert e.value.code == cl._lib.K["CLB_ERR_INVALID"]
        s.close()
    w = workloads.make_workload("richards", 32, N=40, seed=1)
    s = cuda_solver(w, kernel_variant=cl.VARIANT_LANE_PER_CELL)
    with pytest.raises(cl.ClbError):
        s.implicit_step(1800.0, 2)
    s.close()


def test_nan_state_is_reported_not_raised():
    """NaNCheckCallback is a warning in the reference (utils.jl:639-661): a non-finite state comes back in
    clb_stats.nan_count, the call itself succeeds."""
    cl = _cl()
    from climaland_b200 import workloads
    from helpers import cuda_solver
    w = workloads.make_workload("energy_hydrology", 64, N=15, seed=2, topmodel=True)
    w["y_rho_e_int"][5, 3] = np.nan
    s = cuda_solver(w)
    st = s.implicit_step(900.0, 3, want_stats=True)
    assert st["nan_count"] > 0
    s.close()
