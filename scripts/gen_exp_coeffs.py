"""Generate the polynomial coefficients of the branch-free device exp / exp10 kernels in core.cuh.

exp(r)  on |r| <= ln2/2      and   10^r on |r| <= log10(2)/2, degree 11, constrained c0 = 1: Chebyshev-node
interpolation in 60-digit arithmetic (near-minimax), rounded to double.  Prints C initialisers and
the maximum relative error of the rounded polynomial evaluated in double Horner form.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60
DEG = 11


def fit(f, half):
    n = DEG + 1                  # DEG+1 Chebyshev nodes (even count: no node at 0); c0 comes out 1 to 1e-18 and is set to 1
    nodes = [half * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
    sol = mp.lu_solve(A, mp.matrix([f(x) for x in nodes]))
    return [mp.mpf(1)] + [sol[j] for j in range(1, n)]


def check(coef, f, half):
    c = [float(x) for x in coef]
    worst = 0.0
    for r in np.linspace(-half, half, 20001):
        p = c[-1]
        for k in range(len(c) - 2, -1, -1):
            p = p * r + c[k]
        ex = f(mp.mpf(float(r)))
        worst = max(worst, abs(float((mp.mpf(p) - ex) / ex)))
    return worst


for name, f, half in (("EXP_E", mp.exp, mp.log(2) / 2), ("EXP_10", lambda x: mp.power(10, x), mp.log10(2) / 2)):
    coef = fit(f, half)
    print(f"// {name}: max rel err of the double Horner form = {check(coef, f, float(half)):.3e}")
    print("{ " + ", ".join(f"{float(c)!r}" for c in coef) + " }")

ln2 = mp.log(2); hi = float(ln2); lo = float(ln2 - mp.mpf(hi))
print("LN2_HI", repr(hi), "LN2_LO", repr(lo), "L2E", repr(float(1 / ln2)))
lg2 = mp.log10(2); hi = float(lg2); lo = float(lg2 - mp.mpf(hi))
print("LG2_HI", repr(hi), "LG2_LO", repr(lo), "L2_10", repr(float(1 / lg2)))
print("LN10", repr(float(mp.log(10))))
