"""The end-to-end entry point clb_implicit_step_host (host buffers in the reference layout in, new state
out): the pipelined route (column chunks, H2D / kernels / D2H overlapped on three streams) and the
field-by-field route must both give the oracle's stage, and each other's bits."""
import os

import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem

pytestmark = pytest.mark.gpu
TOL = 1e-12
IN_EH = ("y_theta_l", "y_rho_e_int", "y_theta_i", "k_lag", "kappa_lag", "theta_l_lag", "is_saturated", "top_bc_w",
         "bot_bc_w", "top_bc_h", "bot_bc_h", "r_ss", "r_ess", "h_grad", "y_intf_w", "y_intf_e")
IN_RI = ("y_theta_l", "is_saturated", "top_bc_w", "bot_bc_w", "r_ss", "h_grad", "y_intf_w")


def _pin(a):
    """a pinned (page-locked) copy: the library then reads / writes the host array directly (zero-copy route)"""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    t.numpy()[...] = a
    _pin.keep.append(t)
    return t.numpy()


_pin.keep = []


def _run(model, ncol, out_of_place, seed=3, steps=1, pinned=False, options=None):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    eh = model == "energy_hydrology"
    dt, iters = (900.0, 3) if eh else (1800.0, 2)
    w = workloads.make_workload(model, ncol, N=15, seed=seed, topmodel=True)
    P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1)
    s = cuda_solver(w, out_of_place=out_of_place)
    for k, v in (options or {}).items():
        s.set_option(k, v)
    names = IN_EH if eh else IN_RI
    pre = "u" if out_of_place else "y"
    outs = {f"{pre}_theta_l": np.zeros((ncol, 15)), f"{pre}_intf_w": np.zeros(ncol)}
    if eh:
        outs.update({f"{pre}_rho_e_int": np.zeros((ncol, 15)), f"{pre}_intf_e": np.zeros(ncol)})
    ins = {k: np.ascontiguousarray(w[k]) for k in names}
    if pinned:
        ins = {k: _pin(v) for k, v in ins.items()}
        outs = {k: _pin(v) for k, v in outs.items()}
    for _ in range(steps):
        P.implicit_step(U, dt, iters, p=p)
        s.implicit_step_host(dt, iters, ins, outs)
        if steps > 1:  # feed the new state back, as a time loop does
            ins["y_theta_l"][...] = outs[f"{pre}_theta_l"]
            ins["y_intf_w"][...] = outs[f"{pre}_intf_w"]
            if eh:
                ins["y_rho_e_int"][...] = outs[f"{pre}_rho_e_int"]
                ins["y_intf_e"][...] = outs[f"{pre}_intf_e"]
    variant = s.last_variant()
    s.close()
    return U, outs, pre, variant


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("out_of_place", [True, False], ids=["outofplace", "inplace"])
@pytest.mark.parametrize("model,ncol", [("energy_hydrology", 61206), ("energy_hydrology", 9001), ("richards", 20000),
                                        ("energy_hydrology", 700)])
def test_host_step_matches_oracle(model, ncol, out_of_place, pinned):
    """pageable arrays take the staged route (H2D copies into a device staging area), pinned arrays the zero-copy
    route (the relayout kernels address the host arrays directly)"""
    U, outs, pre, _ = _run(model, ncol, out_of_place, pinned=pinned)
    assert_close(outs[f"{pre}_theta_l"], U.theta_l, TOL, "theta_l")
    assert_close(outs[f"{pre}_intf_w"], U.intF_w, TOL, "intF_w")
    if model == "energy_hydrology":
        assert_close(outs[f"{pre}_rho_e_int"], U.rho_e_int, TOL, "rho_e_int")
        assert_close(outs[f"{pre}_intf_e"], U.intF_e, TOL, "intF_e")


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_host_step_repeated_calls(pinned):
    """three stages in a row through the pipelined route (staging buffers and events are reused)"""
    U, outs, pre, variant = _run("energy_hydrology", 30000, True, steps=3, pinned=pinned)
    assert variant == 5
    assert_close(outs["u_theta_l"], U.theta_l, 1e-11, "theta_l after 3 stages")
    assert_close(outs["u_rho_e_int"], U.rho_e_int, 1e-11, "rho_e_int after 3 stages")


def test_pipelined_and_plain_routes_agree_bitwise():
    """CLB_OPT_HOST_ROUTE = 2 forces the field-by-field route, 1 the copy engines + staging, 3 the zero-copy kernels (0: the
    library's choice between those two, by what the host favours), CLB_OPT_TILE_BOXES
    = 1 the lane kernel's one-box-per-field tile requests, CLB_OPT_HOST_CHUNKS another chunking: same bits."""
    res = {}
    for tag, opt in (("auto", {}), ("zerocopy", {"host_route": 3}), ("staged", {"host_route": 1}), ("plain", {"host_route": 2}),
                     ("boxes_per_field", {"tile_boxes": 1}), ("chunks7", {"host_chunks": 7})):
        _, outs, _, _ = _run("energy_hydrology", 20000, True, pinned=True, options=opt)
        res[tag] = outs["u_theta_l"].copy()
    for tag in ("auto", "staged", "zerocopy", "boxes_per_field", "chunks7"):
        assert np.array_equal(res[tag], res["plain"]), tag


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("whole_step", [False, True], ids=["stage", "whole_step"])
def test_host_step_with_a_land_sea_mask(whole_step, pinned):
    """mask_test.jl:53-61 through the host-buffer entry points: a ~1 degree lat-long domain (64 800 columns) of which 30 %
    are land -- the handle holds the active columns only, the caller's arrays the whole domain.  Pinned arrays take the
    chunked zero-copy route (chunks of ACTIVE columns, the host side of every access through the index), pageable ones
    the field-by-field route: both give the compacted problem's stage (or whole step) in the active rows and leave the
    others bit for bit."""
    import climaland_b200 as cl
    from climaland_b200 import workloads
    from test_cuda_soil_step_host import RUNOFF, _oracle_step
    ncol_total, N, dt, iters = 64800, 15, 900.0, 3
    rng = np.random.default_rng(12)
    active = np.sort(rng.choice(ncol_total, int(0.3 * ncol_total), replace=False)).astype(np.int64)
    w = workloads.make_workload("energy_hydrology", ncol_total, N=N, seed=23, topmodel=True)
    xp = workloads.make_explicit_params(w, 23)
    forcing = dict(precip=-rng.uniform(0.0, 4e-7, ncol_total), f_max=rng.uniform(0.2, 0.6, ncol_total))
    s = cl.SoilColumnSolver(model=cl.ENERGY_HYDROLOGY, n_columns=active.size, n_columns_total=ncol_total, z_f=w["z_f"],
                            z_c=w["z_c"], active_columns=active, has_topmodel_source=True)
    for k, v in {**w, **xp}.items():
        if k.lower() in cl.FIELDS:
            s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    s.set("f_max", forcing["f_max"])
    s.set_runoff_params(**RUNOFF)
    # the compacted problem on the oracle
    sub = {k: (v[active] if isinstance(v, np.ndarray) and v.shape[:1] == (ncol_total,) else v) for k, v in w.items()}
    sub["ncol"] = active.size
    P, U, p = oracle_problem(sub, nthreads=os.cpu_count() or 1)
    if whole_step:
        X = P.explicit_params(**{k: v[active] for k, v in xp.items()})
        _oracle_step(P, U, p, X, {k: v[active] for k, v in forcing.items()}, dt, iters)
        ins = {k: np.ascontiguousarray(w[k]).copy() for k in ("y_theta_l", "y_rho_e_int", "y_theta_i", "y_intf_w", "y_intf_e")}
        ins["precip"] = forcing["precip"].copy()
    else:
        P.implicit_step(U, dt, iters, p=p)
        ins = {k: np.ascontiguousarray(w[k]).copy() for k in IN_EH}
    sentinel = -777.0
    outs = {"y_theta_l": np.full((ncol_total, N), sentinel), "y_rho_e_int": np.full((ncol_total, N), sentinel),
            "y_intf_w": np.full(ncol_total, sentinel), "y_intf_e": np.full(ncol_total, sentinel)}
    if pinned:
        ins = {k: _pin(v) for k, v in ins.items()}
        outs = {k: _pin(v) for k, v in outs.items()}
    if whole_step:
        s.soil_step_host(dt, iters, ins, outs)
    else:
        s.implicit_step_host(dt, iters, ins, outs)
    inactive = np.setdiff1d(np.arange(ncol_total), active)
    for k, want in (("y_theta_l", U.theta_l), ("y_rho_e_int", U.rho_e_int), ("y_intf_w", U.intF_w), ("y_intf_e", U.intF_e)):
        assert_close(outs[k][active], want, TOL, k[2:])
        assert np.all(outs[k][inactive] == sentinel), k
    s.close()
