"""Generates the polynomial coefficients used by climaland.jl_b200/csrc/soil_math.cuh.
exp: Chebyshev-node interpolant of exp(r) on |r| <= ln2/2, degree 11, monomial basis.
Run: python tools/gen_math_coeffs.py"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_interp_monomial(f, a, deg):
    n = deg + 1
    nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    V = mp.matrix(n, n)
    for i, x in enumerate(nodes):
        for j in range(n):
            V[i, j] = x ** j
    y = mp.matrix([f(x) for x in nodes])
    return [mp.mpf(c) for c in mp.lu_solve(V, y)]


def horner64(c, x):
    acc = np.full_like(x, c[-1])
    for ck in c[-2::-1]:
        acc = acc * x + ck
    return acc


a = mp.log(2) / 2 * mp.mpf("1.0001")
c = cheb_interp_monomial(mp.exp, a, 11)
c64 = [float(ci) for ci in c]
x = np.linspace(-float(a), float(a), 200001)
approx = horner64(c64, x)
exact = np.array([float(mp.exp(mp.mpf(float(v)))) for v in x[::50]])
print("exp deg 11: max rel err", np.max(np.abs(approx[::50] / exact - 1.0)))
for i, ci in enumerate(c64):
    print(f"    {ci!r},  // r^{i}   ({ci.hex()})")
