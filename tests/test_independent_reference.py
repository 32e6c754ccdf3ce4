"""Rows a15 / a16 (linear solve, Newton / ARS111 stage) pinned to the MATHEMATICS by a second, independent statement:
tests/independent_reference.py (extended precision, operator matrices, dense Gaussian elimination of the full block
system) against the C oracle (CPU, here) and against the CUDA library (GPU).  The two share no code.

Tolerances, written out: the stage's new state agrees to 1e-12 element-wise relative (floor: 1e-6 of the field's
largest magnitude -- rho_e_int crosses zero); the solve alone (`ldiv!` of a given right-hand side against the dense
solution of the assembled matrix) to max(1e-12, 4 cond(W) 2^-53) of the solution's largest entry per column and block
(the forward-error bound of any double-precision solve against the exact solution)."""
import numpy as np
import pytest

import independent_reference as ind
from helpers import cuda_solver, elem_rel_err, oracle_problem
from climaland_b200 import workloads

CASES = [("richards", 0, False, 2, 1800.0), ("richards", 1, True, 2, 1800.0),
         ("energy_hydrology", 0, True, 3, 900.0), ("energy_hydrology", 0, False, 3, 900.0),
         ("energy_hydrology", 1, True, 3, 900.0)]
IDS = [f"{m}-cl{c}-tm{int(t)}" for m, c, t, _, _ in CASES]
NCOL = 12


def _workload(model, closure, topmodel, N=15, seed=5):
    w = workloads.make_workload(model, NCOL, N=N, seed=seed, topmodel=topmodel)
    if closure == 1:  # Brooks-Corey: c in the slot of alpha, psi_b in the slot of n; state consistent with them
        from test_cuda_hooks_parity import _to_brooks_corey
        w = _to_brooks_corey(w)
    return w


def _independent_step(w, closure, dt, iters):
    out = {k: np.zeros_like(w["y_theta_l"]) for k in ("theta_l", "rho_e_int")}
    intf = {k: np.zeros(w["ncol"]) for k in ("intF_w", "intF_e")}
    N = w["N"]
    for c in range(w["ncol"]):
        x = ind.Column(w, c, closure=closure).implicit_step(dt, iters)
        out["theta_l"][c] = x[:N].astype(np.float64)
        if w["model"] == "energy_hydrology":
            out["rho_e_int"][c] = x[N:2 * N].astype(np.float64)
            intf["intF_w"][c], intf["intF_e"][c] = float(x[2 * N]), float(x[2 * N + 1])
        else:
            intf["intF_w"][c] = float(x[N])
    return out, intf


def _check(got, ref, model, what):
    names = ("theta_l", "rho_e_int") if model == "energy_hydrology" else ("theta_l",)
    for n in names:
        e = elem_rel_err(got[n], ref[n], floor_rel=1e-6)
        assert e <= 1e-12, f"{what}: {n} element-wise relative error {e:.3e} > 1e-12"


@pytest.mark.parametrize("model,closure,topmodel,iters,dt", CASES, ids=IDS)
def test_oracle_stage_equals_the_independent_dense_reference(model, closure, topmodel, iters, dt):
    w = _workload(model, closure, topmodel)
    ref, rint = _independent_step(w, closure, dt, iters)
    P, U, p = oracle_problem(w, closure=closure)
    P.implicit_step(U, dt, iters, p=p)
    _check(dict(theta_l=U.theta_l, rho_e_int=U.rho_e_int), ref, model, "oracle stage")
    assert elem_rel_err(U.intF_w, rint["intF_w"], floor_rel=1e-6) <= 1e-12
    if model == "energy_hydrology":
        assert elem_rel_err(U.intF_e, rint["intF_e"], floor_rel=1e-6) <= 1e-12


@pytest.mark.parametrize("model", ["richards", "energy_hydrology"])
def test_oracle_ldiv_equals_the_dense_solution_of_the_assembled_matrix(model):
    """x = ldiv!(W, b) (Thomas per block, block lower triangular, x = W^-1 b): the oracle's solve against Gaussian
    elimination of the dense matrix the independent reference assembles from operator matrices."""
    w = _workload(model, 0, True)
    N, dtg = w["N"], 900.0
    P, U, p = oracle_problem(w)
    P.update_implicit_cache(U, p)
    W = P.new_jacobian()
    P.compute_jacobian(W, U, p, dtg)
    rng = np.random.default_rng(8)
    b = P.new_state()
    b.theta_l[...] = rng.normal(0, 1e-3, b.theta_l.shape)
    b.intF_w[...] = rng.normal(0, 1e-6, w["ncol"])
    if model == "energy_hydrology":
        b.rho_e_int[...] = rng.normal(0, 1e4, b.theta_l.shape)
        b.intF_e[...] = rng.normal(0, 1.0, w["ncol"])
    x = P.new_state()
    P.ldiv(x, W, b)
    eh = model == "energy_hydrology"

    def tri(blk, c):  # dense N x N block from the oracle's own entries: lo[i] = (i, i-1), up[i] = (i, i+1)
        lo, di, up = (getattr(W, f"w{blk}_{d}")[c].astype(ind.LD) for d in ("lo", "di", "up"))
        return np.diag(di) + np.diag(lo[1:], -1) + np.diag(up[:-1], 1)

    for c in range(w["ncol"]):
        # (1) the SOLVE alone: the oracle's matrix entries, taken as exact numbers, solved densely in extended precision
        n = (2 * N + 2) if eh else (N + 1)
        Wd = np.zeros((n, n), dtype=ind.LD)
        Wd[:N, :N] = tri("11", c)
        if eh:
            Wd[N:2 * N, :N], Wd[N:2 * N, N:2 * N] = tri("21", c), tri("22", c)
            Wd[2 * N, 2 * N] = Wd[2 * N + 1, 2 * N + 1] = -1
            bb = np.concatenate([b.theta_l[c], b.rho_e_int[c], [b.intF_w[c], b.intF_e[c]]]).astype(ind.LD)
        else:
            Wd[N, N] = -1
            bb = np.concatenate([b.theta_l[c], [b.intF_w[c]]]).astype(ind.LD)
        xd = ind.dense_solve(Wd, bb).astype(np.float64)
        got = [x.theta_l[c]] + ([x.rho_e_int[c]] if eh else [])
        # forward error of a double-precision solve against the exact solution of the same matrix: <= ~cond 2^-53;
        # stated bound max(1e-12, 8 cond_inf(diagonal block) 2^-53) of the block's largest entry
        for k, g in enumerate(got):
            blk = Wd[k * N:(k + 1) * N, k * N:(k + 1) * N].astype(np.float64)
            bound = max(1e-12, 8.0 * np.linalg.cond(blk, np.inf) * 2.0 ** -53)
            assert bound < 1e-10, bound
            r = xd[k * N:(k + 1) * N]
            assert np.max(np.abs(g - r)) <= bound * np.max(np.abs(r)), (c, k, bound)
        iw = 2 * N if eh else N
        assert abs(x.intF_w[c] - xd[iw]) <= 1e-15 * abs(xd[iw]) + 1e-300
        # (2) the MATRIX: the oracle's entries against the operator-matrix products of the independent reference.
        # dpsi/dtheta loses digits near saturation in ANY double evaluation of the reference's formula (S^(-1/m) - 1
        # cancels), so the entries are held to 1e-9 of the block's largest entry here; the reference-held Jacobian
        # known answers (tests/test_oracle_reference_kats.py) pin them at the reference's own tolerance.
        col = ind.Column(w, c)
        xs = col.state()
        _, Wi = col.residual_and_jacobian(xs, xs, ind.LD(dtg))
        for (r0, c0) in ([(0, 0), (N, 0), (N, N)] if eh else [(0, 0)]):
            A, B = Wd[r0:r0 + N, c0:c0 + N].astype(np.float64), Wi[r0:r0 + N, c0:c0 + N].astype(np.float64)
            assert np.max(np.abs(A - B)) <= 1e-9 * np.max(np.abs(B)), (c, r0, c0)


@pytest.mark.gpu
@pytest.mark.parametrize("model,closure,topmodel,iters,dt", CASES, ids=IDS)
def test_cuda_stage_equals_the_independent_dense_reference(model, closure, topmodel, iters, dt):
    w = _workload(model, closure, topmodel)
    ref, rint = _independent_step(w, closure, dt, iters)
    s = cuda_solver(w, closure=closure)
    s.implicit_step(dt, iters)
    got = dict(theta_l=s.get("y_theta_l"))
    if model == "energy_hydrology":
        got["rho_e_int"] = s.get("y_rho_e_int")
    _check(got, ref, model, f"CUDA stage (variant {s.last_variant()})")
    assert elem_rel_err(s.get("y_intf_w"), rint["intF_w"], floor_rel=1e-6) <= 1e-12
    s.close()
