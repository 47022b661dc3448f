"""1-D node sets, barycentric weights and Lagrange polynomials (host side).

These are the small host-side tables the Fortran host hands to the device layer at
init (`N_Inter(N)%xGP/wGP/wBary`, `XiCL_NGeo`, `wBaryCL_NGeo`).  They are computed
with the same algorithms the reference uses so that a stand-alone run (tests, bench)
feeds the kernels the same numbers a PICLas host would:

* Legendre-Gauss nodes/weights  : reference src/interpolation/basis.f90:758-833
  (Newton iteration on L_{N+1} started from Chebyshev points, tol 1e-15, <=10 its)
* Legendre polynomial + derivative: basis.f90:643-688
* Chebyshev-Gauss-Lobatto nodes : basis.f90:719-754
* barycentric weights           : basis.f90:952-976
* Lagrange basis at a point     : basis.f90:1223-1264 (with the node-hit branch,
  ALMOSTEQUAL_UNITY basis.f90:1011-1035)
"""
from __future__ import annotations

import math
import numpy as np

PP_REAL_TOLERANCE = np.finfo(np.float64).eps  # preprocessing.f90:26


def legendre_poly_and_derivative(n: int, x: float):
    if n == 0:
        L, Lder = 1.0, 0.0
    elif n == 1:
        L, Lder = x, 1.0
    else:
        L_Nm2, L_Nm1 = 1.0, x
        Lder_Nm2, Lder_Nm1 = 0.0, 1.0
        L = Lder = 0.0
        for i in range(2, n + 1):
            L = (float(2 * i - 1) * x * L_Nm1 - float(i - 1) * L_Nm2) / float(i)
            Lder = Lder_Nm2 + float(2 * i - 1) * L_Nm1
            L_Nm2, L_Nm1 = L_Nm1, L
            Lder_Nm2, Lder_Nm1 = Lder_Nm1, Lder
    s = math.sqrt(float(n) + 0.5)
    return L * s, Lder * s


def legendre_gauss_nodes_weights(n: int):
    x = np.zeros(n + 1)
    w = np.zeros(n + 1)
    if n == 0:
        x[:] = 0.0
        w[:] = 2.0
        return x, w
    if n == 1:
        x[0] = -math.sqrt(1.0 / 3.0)
        x[1] = -x[0]
        w[:] = 1.0
        return x, w
    tol = 1.0e-15
    cheb = 2.0 * math.atan(1.0) / float(n + 1)
    for i in range((n + 1) // 2):
        xi = -math.cos(cheb * float(2 * i + 1))
        for _ in range(11):
            L, Ld = legendre_poly_and_derivative(n + 1, xi)
            dx = -L / Ld
            xi += dx
            if abs(dx) < tol * abs(xi):
                break
        L, Ld = legendre_poly_and_derivative(n + 1, xi)
        x[i] = xi
        x[n - i] = -xi
        w[i] = (2.0 * n + 3) / ((1.0 - xi * xi) * Ld * Ld)
        w[n - i] = w[i]
    if n % 2 == 0:
        x[n // 2] = 0.0
        L, Ld = legendre_poly_and_derivative(n + 1, 0.0)
        w[n // 2] = (2.0 * n + 3) / (Ld * Ld)
    return x, w


def cheb_gauss_lobatto_nodes(n: int):
    if n == 0:
        return np.zeros(1)
    return np.array([-math.cos(i / float(n) * math.acos(-1.0)) for i in range(n + 1)])


def barycentric_weights(x: np.ndarray):
    n = len(x) - 1
    w = np.ones(n + 1)
    for i in range(1, n + 1):
        for j in range(i):
            w[j] = w[j] * (x[j] - x[i])
            w[i] = w[i] * (x[i] - x[j])
    return 1.0 / w


def almost_equal_unity(x: float, y: float) -> bool:
    tol = PP_REAL_TOLERANCE
    if x == 0.0 or y == 0.0:
        return abs(x - y) <= 2.0 * tol
    return abs(x - y) <= tol * abs(x) and abs(x - y) <= tol * abs(y)


def lagrange_polys(x: float, xgp: np.ndarray, wbary: np.ndarray):
    n = len(xgp) - 1
    L = np.zeros(n + 1)
    hit = False
    for i in range(n + 1):
        if almost_equal_unity(x, xgp[i]):
            L[i] = 1.0
            hit = True
    if hit:
        return L
    s = 0.0
    for i in range(n + 1):
        L[i] = wbary[i] / (x - xgp[i])
        s += L[i]
    return L / s


def poly_derivative_matrix(x: np.ndarray):
    """D_ij = l_j'(x_i) for the nodal basis on x (basis.f90 PolynomialDerivativeMatrix)."""
    n = len(x) - 1
    w = barycentric_weights(x)
    D = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(n + 1):
            if i != j:
                D[i, j] = w[j] / (w[i] * (x[i] - x[j]))
                D[i, i] -= D[i, j]
    return D
