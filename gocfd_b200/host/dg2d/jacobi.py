"""Orthonormal Jacobi / simplex polynomial bases (host side, runs once at start-up).

Mirrors the pieces of the reference that build the constant operators handed to the
device library.  These stay on the host in the reference too (north star: "DG2D
basis/operator construction ... remain in Go").

Reference: DG1D/elements.go:257-329 (JacobiP, GradJacobiP, Vandermonde1D),
DG1D/utils.go:35-46 (Gamma0, Gamma1), DG2D/basis_polynomials.go:55-153
(Vandermonde2D, Simplex2DP, GradSimplex2DP), DG2D/element_utils.go:111-140 (rsToab).
"""
import math

import numpy as np


def _gamma0(alpha, beta):
    ab1 = alpha + beta + 1.0
    return (math.gamma(alpha + 1.0) * math.gamma(beta + 1.0)
            * math.pow(2.0, ab1) / ab1 / math.gamma(ab1))


def _gamma1(alpha, beta):
    return (alpha + 1.0) * (beta + 1.0) * _gamma0(alpha, beta) / (alpha + beta + 3.0)


def jacobi_p(r, alpha, beta, n):
    """Orthonormal Jacobi polynomial P_n^(alpha,beta) at points r (three-term recurrence)."""
    r = np.asarray(r, dtype=np.float64)
    p_prev = np.full(r.shape, 1.0 / math.sqrt(_gamma0(alpha, beta)))
    if n == 0:
        return p_prev
    ab = alpha + beta
    p_cur = (1.0 / math.sqrt(_gamma1(alpha, beta))) * ((ab + 2.0) * r / 2.0 + (alpha - beta) / 2.0)
    if n == 1:
        return p_cur
    a1, b1, ab1 = alpha + 1.0, beta + 1.0, ab + 1.0
    aold = 2.0 * math.sqrt(a1 * b1 / (ab + 3.0)) / (ab + 2.0)
    for i in range(n - 1):
        ip1 = float(i + 1)
        ip2 = ip1 + 1.0
        h1 = 2.0 * ip1 + ab
        anew = 2.0 / (h1 + 2.0) * math.sqrt(
            ip2 * (ip1 + ab1) * (ip1 + a1) * (ip1 + b1) / (h1 + 1.0) / (h1 + 3.0))
        bnew = -(alpha * alpha - beta * beta) / h1 / (h1 + 2.0)
        p_next = (-aold * p_prev + (r - bnew) * p_cur) / anew
        p_prev, p_cur = p_cur, p_next
        aold = anew
    return p_cur


def grad_jacobi_p(r, alpha, beta, n):
    r = np.asarray(r, dtype=np.float64)
    if n == 0:
        return np.zeros(r.shape)
    return jacobi_p(r, alpha + 1.0, beta + 1.0, n - 1) * math.sqrt(n * (n + alpha + beta + 1.0))


def vandermonde_1d(n, r):
    r = np.asarray(r, dtype=np.float64)
    return np.stack([jacobi_p(r, 0.0, 0.0, j) for j in range(n + 1)], axis=1)


def grad_vandermonde_1d(n, r):
    r = np.asarray(r, dtype=np.float64)
    return np.stack([grad_jacobi_p(r, 0.0, 0.0, j) for j in range(n + 1)], axis=1)


def rs_to_ab(r, s):
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    safe = np.where(s != 1.0, 1.0 - s, 1.0)
    a = np.where(s != 1.0, 2.0 * (1.0 + r) / safe - 1.0, -1.0)
    return a, s.copy()


def simplex_2d_p(r, s, i, j):
    """sqrt(2) P_i^(0,0)(a) P_j^(2i+1,0)(b) (1-b)^i, the (i,j) orthonormal mode on the triangle."""
    a, b = rs_to_ab(r, s)
    h1 = jacobi_p(a, 0.0, 0.0, i)
    h2 = jacobi_p(b, 2.0 * i + 1.0, 0.0, j)
    return math.sqrt(2.0) * h1 * h2 * (1.0 - b) ** i


def grad_simplex_2d_p(r, s, i, j):
    a, b = rs_to_ab(r, s)
    fa = jacobi_p(a, 0.0, 0.0, i)
    dfa = grad_jacobi_p(a, 0.0, 0.0, i)
    gb = jacobi_p(b, 2.0 * i + 1.0, 0.0, j)
    dgb = grad_jacobi_p(b, 2.0 * i + 1.0, 0.0, j)
    half = 0.5 * (1.0 - b)
    norm = math.pow(2.0, i + 0.5)
    ddr = dfa * gb
    if i > 0:
        ddr = ddr * half ** (i - 1)
    ddr = ddr * norm
    dds = 0.5 * dfa * gb * (1.0 + a)
    if i > 0:
        dds = dds * half ** (i - 1)
    tmp = dgb * half ** i
    if i > 0:
        tmp = tmp - 0.5 * i * gb * half ** (i - 1)
    dds = (dds + fa * tmp) * norm
    return ddr, dds


def mode_orders(n):
    """(i, j) of every mode in column order: for i in 0..N, for j in 0..N-i."""
    return [(i, j) for i in range(n + 1) for j in range(n + 1 - i)]


def vandermonde_2d(n, r, s):
    return np.stack([simplex_2d_p(r, s, i, j) for (i, j) in mode_orders(n)], axis=1)


def grad_vandermonde_2d(n, r, s):
    cols = [grad_simplex_2d_p(r, s, i, j) for (i, j) in mode_orders(n)]
    return (np.stack([c[0] for c in cols], axis=1), np.stack([c[1] for c in cols], axis=1))


def inverse_with_check(m):
    """Dense inverse plus the reference's sanity check (utils/matrix_extended.go:1323-1346)."""
    inv = np.linalg.inv(m)
    total = float((m @ inv).sum())
    if abs(total - m.shape[0]) > 1e-6:
        raise ArithmeticError("inversion of Vandermonde matrix failed: sum %.3f, expected %d"
                              % (total, m.shape[0]))
    return inv
