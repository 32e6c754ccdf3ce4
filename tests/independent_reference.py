"""An INDEPENDENT second statement of the implicit soil stage, for a handful of columns, in extended precision.

Purpose ("narrow the unpinned part of parity"): the linear solve (ClimaCore field_matrix_solve!) and
the Newton / ARS111 stage (ClimaTimeSteppers) are not vendored in /root/reference and there is no Julia here, so
`oracle/soil_oracle.c` restates their published algorithms.  This module shares NO code with that oracle and takes a
different route to the same mathematics, so that an error in the oracle's Thomas ordering, block elimination or Newton
update cannot hide:

  * numpy `longdouble` throughout (x86 80-bit: 64-bit mantissa, ~1e-19), not double;
  * the finite-difference operators are explicit MATRICES composed as the reference composes its operator matrices
    (energy_hydrology.jl:466-576, rre.jl:391-458): D = DivergenceF2C (N x N+1), G = GradientC2F with zero boundary
    gradient (N+1 x N), I_f = InterpolateC2F with extrapolated boundaries (N+1 x N); W = -dtgamma D diag(I_f(-K)) G
    diag(dpsi) - I etc. by matrix products, not by a stencil;
  * the Newton update solves the FULL dense system [[W11, 0], [W21, W22]] (and the -I flux-integral rows) by Gaussian
    elimination with partial pivoting, not by Thomas sweeps and not block by block: what
    implicit_timestepping.jl:160-171 (BlockLowerTriangularSolve) computes must equal it;
  * the stage is ClimaTimeSteppers' Newton iteration literally: x <- x - W(x)^-1 (temp + dtgamma T_imp(x) - x), the
    Jacobian and the implicit cache re-evaluated at every iterate (Simulations.jl:127-135).

Closures follow soil_hydrology_parameterizations.jl:45-173, 220-289 (van Genuchten / Brooks-Corey), the temperature
soil_heat_parameterizations.jl:137-212, the TOPMODEL implicit source Runoff.jl:321-359.  Only flux boundary values
(the lagged top / bottom fluxes) are covered: that is the bench configuration.

Test infrastructure only (tests/ import it); pure numpy, slow, a few columns.
"""
import numpy as np

LD = np.longdouble
SQRT_EPS = LD(np.sqrt(np.finfo(np.float64).eps))
EPS = LD(np.finfo(np.float64).eps)


def _ld(a):
    return np.asarray(a, dtype=LD)


# ---- closures (per cell arrays, longdouble) ---------------------------------------------------------------------
def effective_saturation(nu, theta, theta_r):
    th_safe = np.maximum(theta, theta_r + SQRT_EPS)
    nu_safe = np.maximum(nu, theta_r + SQRT_EPS)
    return (th_safe - theta_r) / (nu_safe - theta_r), th_safe, nu_safe


def vg_psi_dpsi_K(theta, nu_eff, theta_r, alpha, n, m, S_s, K_sat):
    S, th_safe, nu_safe = effective_saturation(nu_eff, theta, theta_r)
    Sc = np.minimum(S, LD(1))  # evaluated everywhere, selected below
    with np.errstate(all="ignore"):
        psi_u = -((Sc ** (-1 / m) - 1) * alpha ** (-n)) ** (1 / n)
        dpsi_u = 1 / (alpha * m * n) / (nu_safe - theta_r) * (Sc ** (-1 / m) - 1) ** (1 / n - 1) * Sc ** (-1 / m - 1)
        K_u = K_sat * np.sqrt(Sc) * (1 - (1 - Sc ** (1 / m)) ** m) ** 2
    psi = np.where(S <= 1, psi_u, (th_safe - nu_safe) / S_s)
    dpsi = np.where(S < 1, dpsi_u, 1 / S_s)
    K = np.where(S < 1, K_u, K_sat)
    return psi, dpsi, K


def bc_psi_dpsi_K(theta, nu_eff, theta_r, psi_b, c, S_s, K_sat):
    S, th_safe, nu_safe = effective_saturation(nu_eff, theta, theta_r)
    Sc = np.minimum(S, LD(1))
    psi = np.where(S <= 1, psi_b * Sc ** (-1 / c), (th_safe - nu_safe) / S_s + psi_b)
    dpsi = np.where(S < 1, -psi_b / (c * (nu_safe - theta_r)) * Sc ** (-(1 + 1 / c)), 1 / S_s)
    K = np.where(S < 1, K_sat * Sc ** (2 / c + 3), K_sat)
    return psi, dpsi, K


# ---- operator matrices of one column (z_c, z_f: level 0 = bottom) ------------------------------------------------
def operators(z_c, z_f):
    N = len(z_c)
    dz_c = np.diff(_ld(z_f))
    D = np.zeros((N, N + 1), dtype=LD)  # (div q)_i = (q_{i+1/2} - q_{i-1/2}) / dz_i
    for i in range(N):
        D[i, i], D[i, i + 1] = -1 / dz_c[i], 1 / dz_c[i]
    G = np.zeros((N + 1, N), dtype=LD)  # interior faces only; boundary rows zero (SetGradient(0) / SetValue fluxes)
    for f in range(1, N):
        d = _ld(z_c)[f] - _ld(z_c)[f - 1]
        G[f, f - 1], G[f, f] = -1 / d, 1 / d
    If = np.zeros((N + 1, N), dtype=LD)  # arithmetic mean; extrapolated at the two boundaries
    for f in range(1, N):
        If[f, f - 1] = If[f, f] = LD(0.5)
    If[0, 0] = If[N, N - 1] = LD(1)
    return D, G, If


def dense_solve(A, b):
    """Gaussian elimination with partial pivoting, longdouble."""
    A, b = A.copy(), b.copy()
    n = len(b)
    for k in range(n):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]], b[[k, p]] = A[[p, k]], b[[p, k]]
        for i in range(k + 1, n):
            if A[i, k] != 0:
                f = A[i, k] / A[k, k]
                A[i, k:] -= f * A[k, k:]
                b[i] -= f * b[k]
    x = np.zeros(n, dtype=LD)
    for k in range(n - 1, -1, -1):
        x[k] = (b[k] - A[k, k + 1:] @ x[k + 1:]) / A[k, k]
    return x


class Column:
    """One column of a `workloads.make_workload` dict (flux boundary values)."""

    def __init__(self, w, c, closure=0, earth=None):
        from climaland_b200 import workloads
        self.eh = w["model"] == "energy_hydrology"
        self.closure = closure
        self.E = {k: LD(v) for k, v in (earth or workloads.EARTH).items()}
        self.z_c, self.z_f = _ld(w["z_c"]), _ld(w["z_f"])
        self.N = len(self.z_c)
        self.D, self.G, self.If = operators(w["z_c"], w["z_f"])
        g = lambda k: _ld(w[k][c])
        self.nu, self.theta_r, self.K_sat, self.S_s = g("nu"), g("theta_r"), g("K_sat"), g("S_s")
        self.a, self.b = g("hcm_a"), g("hcm_b")
        self.m = g("hcm_m") if closure == 0 else None
        self.top_w, self.bot_w = g("top_bc_w"), g("bot_bc_w")
        self.topmodel = bool(w.get("topmodel", False))
        if self.topmodel:
            self.sat, self.R_ss, self.h_grad = g("is_saturated"), g("r_ss"), g("h_grad")
        self.theta = g("y_theta_l")
        self.intF_w = g("y_intf_w")
        if self.eh:
            self.theta_i, self.rho_e, self.rho_c_ds = g("y_theta_i"), g("y_rho_e_int"), g("rho_c_ds")
            self.K_lag, self.kappa, self.theta_l_lag = g("k_lag"), g("kappa_lag"), g("theta_l_lag")
            self.top_h, self.bot_h, self.intF_e = g("top_bc_h"), g("bot_bc_h"), g("y_intf_e")
            if self.topmodel:
                self.R_ess = g("r_ess")

    def _closure(self, theta, nu_eff):
        if self.closure == 0:
            return vg_psi_dpsi_K(theta, nu_eff, self.theta_r, self.a, self.b, self.m, self.S_s, self.K_sat)
        return bc_psi_dpsi_K(theta, nu_eff, self.theta_r, self.b, self.a, self.S_s, self.K_sat)  # fields: a = c, b = psi_b

    def _flux_div(self, q_int, top, bot):
        """-(D q) with the boundary faces of q set to the boundary fluxes (DivergenceF2C with SetValue)."""
        q = q_int.copy()
        q[0], q[self.N] = bot, top
        return -(self.D @ q)

    def residual_and_jacobian(self, x, temp, dtg):
        """x = [theta (N), rho_e (N), intF_w, intF_e] (EnergyHydrology) or [theta (N), intF_w] (Richards)."""
        N, E = self.N, self.E
        theta = x[:N]
        if self.eh:
            rho_e = x[N:2 * N]
            nu_eff = self.nu - self.theta_i
            theta_l = np.minimum(nu_eff, theta)
            rho_c = self.rho_c_ds + theta_l * E["rho_l"] * E["cp_l"] + self.theta_i * E["rho_i"] * E["cp_i"]
            T = E["T_ref"] + (rho_e + self.theta_i * E["rho_i"] * E["LH_f0"]) / rho_c
            psi, dpsi, _ = self._closure(theta, nu_eff)
            K = self.K_lag
        else:
            psi, dpsi, K = self._closure(theta, self.nu)
        h = psi + self.z_c
        src_w = src_e = LD(0)
        if self.topmodel:
            src_w = self.R_ss / max(self.h_grad, EPS)
            if self.eh:
                src_e = self.R_ess / max(self.h_grad, EPS)
        qw = -(self.If @ K) * (self.G @ h)
        T_theta = self._flux_div(qw, self.top_w, self.bot_w)
        if self.topmodel:
            T_theta = T_theta - src_w * self.sat
        T_intw = -(self.top_w - self.bot_w) - (self.R_ss if self.topmodel else LD(0))
        I = np.eye(N, dtype=LD)
        W11 = -dtg * (self.D @ np.diag(self.If @ (-K)) @ self.G @ np.diag(dpsi)) - I
        if not self.eh:
            n = N + 1
            W = np.zeros((n, n), dtype=LD)
            W[:N, :N] = W11
            W[N, N] = -1
            Timp = np.concatenate([T_theta, [T_intw]])
            return temp + dtg * Timp - x, W
        e_l = E["rho_l"] * E["cp_l"] * (T - E["T_ref"])
        qe = -(self.If @ self.kappa) * (self.G @ T) - (self.If @ (e_l * K)) * (self.G @ h)
        T_rhoe = self._flux_div(qe, self.top_h, self.bot_h)
        if self.topmodel:
            T_rhoe = T_rhoe - src_e * self.sat
        T_inte = -(self.top_h - self.bot_h) - (self.R_ess if self.topmodel else LD(0))
        rho_cJ = self.rho_c_ds + self.theta_l_lag * E["rho_l"] * E["cp_l"] + self.theta_i * E["rho_i"] * E["cp_i"]
        W21 = -dtg * (self.D @ np.diag(-(self.If @ (e_l * K))) @ self.G @ np.diag(dpsi)) - I
        W22 = -dtg * (self.D @ np.diag(self.If @ (-self.kappa)) @ self.G @ np.diag(1 / rho_cJ)) - I
        n = 2 * N + 2
        W = np.zeros((n, n), dtype=LD)
        W[:N, :N], W[N:2 * N, :N], W[N:2 * N, N:2 * N] = W11, W21, W22
        W[2 * N, 2 * N] = W[2 * N + 1, 2 * N + 1] = -1
        Timp = np.concatenate([T_theta, T_rhoe, [T_intw, T_inte]])
        return temp + dtg * Timp - x, W

    def state(self):
        if self.eh:
            return np.concatenate([self.theta, self.rho_e, [self.intF_w, self.intF_e]])
        return np.concatenate([self.theta, [self.intF_w]])

    def implicit_step(self, dtg, max_iters):
        """ARS111 implicit stage: Newton's method, Jacobian every iteration; returns the new state vector."""
        temp = self.state()
        x = temp.copy()
        for _ in range(max_iters):
            f, W = self.residual_and_jacobian(x, temp, LD(dtg))
            x = x - dense_solve(W, f)
        return x
