"""Solution (Lagrange, WSJ points) and flux (Raviart-Thomas) reference elements.

Host-side mirror of the reference's element construction.  Only the constant operator
matrices matter downstream: V / Vinv / MassMatrix of the solution element, and Div /
DivInt of the RT element.  Row/column orders are part of the device ABI (SURVEY.md
appendix A) and follow the reference exactly.

Reference: DG2D/lagrange_element.go:54-91, DG2D/basis_polynomials.go:19-53,
DG2D/raviart_thomas_element.go:192-297, :377-400, :429-496,
DG2D/rt_basis_simplex.go:17-273, DG2D/basis_polynomial_construction.go:74-87,
DG2D/nodes2d_williams_shunn_jameson.go:146-169.
"""
import math

import numpy as np

from . import jacobi as jb
from .wsj_table import wsj_barycentric


def wsj_points(order):
    """(R, S, W) of the Williams-Shunn-Jameson points; r = -l1+l2-l3, s = -l1-l2+l3."""
    b = wsj_barycentric(order)
    r = -b[:, 0] + b[:, 1] - b[:, 2]
    s = -b[:, 0] - b[:, 1] + b[:, 2]
    return r, s, b[:, 3].copy()


class JacobiBasis2D:
    """Orthonormal modal basis of order P sampled on nodes (R, S); V, Vinv, mode orders."""

    def __init__(self, p, r, s):
        self.P = p
        self.Np = (p + 1) * (p + 2) // 2
        self.Order2DAtJ = jb.mode_orders(p)
        self.OrderAtJ = [i + j for (i, j) in self.Order2DAtJ]
        self.V = jb.vandermonde_2d(p, r, s)
        self.Vr, self.Vs = jb.grad_vandermonde_2d(p, r, s)
        self.Vinv = jb.inverse_with_check(self.V)

    def interp_matrix(self, r, s):
        """Nodal values -> values at (r, s): Vandermonde(r, s) . Vinv."""
        return jb.vandermonde_2d(self.P, r, s) @ self.Vinv

    def lagrange(self, r, s):
        """Values of every nodal (Lagrange) polynomial at the points, shape (npts, Np)."""
        return jb.vandermonde_2d(self.P, r, s) @ self.Vinv

    def lagrange_grad(self, r, s):
        vr, vs = jb.grad_vandermonde_2d(self.P, r, s)
        return vr @ self.Vinv, vs @ self.Vinv


class LagrangeElement2D:
    def __init__(self, n):
        if n < 0:
            raise ValueError("Polynomial order must be >= 0, have %d" % n)
        self.N = n
        self.Np = (n + 1) * (n + 2) // 2
        self.R, self.S, self.W = wsj_points(n)
        self.JB2D = JacobiBasis2D(n, self.R, self.S)
        v = self.JB2D.V
        self.MassMatrix = v.T @ np.diag(self.W) @ v

    def derivative_matrices(self, r, s):
        vr, vs = jb.grad_vandermonde_2d(self.N, r, s)
        return vr @ self.JB2D.Vinv, vs @ self.JB2D.Vinv


# Optimised edge-point parameters, indexed by RT order (raviart_thomas_element.go:377-400).
_EDGE_POINTS = {
    1: [-0.38490018, 0.38490018],
    2: [-0.028, 0.000, 0.028],
    3: [-0.482, -0.161, 0.161, 0.482],
    4: [-0.586, -0.293, 0.000, 0.293, 0.586],
    5: [-0.634, -0.381, -0.127, 0.127, 0.381, 0.634],
    6: [-0.671, -0.452, -0.219, 0.000, 0.219, 0.452, 0.671],
    7: [-0.808, -0.578, -0.343, -0.113, 0.113, 0.343, 0.578, 0.808],
    8: [-0.692, -0.523, -0.350, -0.176, 0.000, 0.176, 0.350, 0.523, 0.692],
}


class RTElement:
    """Raviart-Thomas element of order P on the reference triangle (simplex RT basis).

    Row order (Np rows): NpInt interior r-DOFs, NpInt interior s-DOFs (same points), then
    NpEdge points on each of edge 0 (s=-1), edge 1 (hypotenuse), edge 2 (r=-1).
    """

    def __init__(self, p):
        if p < 1:
            raise ValueError("P must be greater than or equal to 1")
        self.P = p
        self.Np = (p + 1) * (p + 3)
        self.NpInt = p * (p + 1) // 2
        self.NpEdge = p + 1
        self.RInt, self.SInt, _ = wsj_points(p - 1)
        gq = np.array(_EDGE_POINTS[p], dtype=np.float64)
        gpt = 0.5 * (gq + 1.0)
        r_edge = np.concatenate([gq, 1.0 - 2.0 * gpt, -np.ones(p + 1)])
        s_edge = np.concatenate([-np.ones(p + 1), -1.0 + 2.0 * gpt, -gq])
        self.R = np.concatenate([self.RInt, self.RInt, r_edge])
        self.S = np.concatenate([self.SInt, self.SInt, s_edge])
        self._edge_param = gq
        self._build()

    def edge_locations(self, f):
        return np.asarray(f)[2 * self.NpInt:]

    def _build(self):
        np_, ni, ne = self.Np, self.NpInt, self.NpEdge
        r, s = self.R, self.S
        oosr2 = 0.5 * math.sqrt(2.0)
        dof = np.zeros((np_, 2))
        dof[:ni] = (1.0, 0.0)
        dof[ni:2 * ni] = (0.0, 1.0)
        dof[2 * ni:2 * ni + ne] = (0.0, -1.0)
        dof[2 * ni + ne:2 * ni + 2 * ne] = (oosr2, oosr2)
        dof[2 * ni + 2 * ne:] = (-1.0, 0.0)
        self.DOFVectors = dof

        rp, rm = 0.5 * (r + 1.0), 0.5 * (r - 1.0)
        sp, sm = 0.5 * (s + 1.0), 0.5 * (s - 1.0)
        sr2 = math.sqrt(2.0)

        # Scalar multipliers psi_j and their gradients at every point of the element.
        psi = np.zeros((np_, np_))
        dpsi_r = np.zeros((np_, np_))
        dpsi_s = np.zeros((np_, np_))
        pk = JacobiBasis2D(self.P - 1, self.RInt, self.SInt)
        lag = pk.lagrange(r, s)
        lag_r, lag_s = pk.lagrange_grad(r, s)
        psi[:, :ni] = lag
        psi[:, ni:2 * ni] = lag
        dpsi_r[:, :ni], dpsi_s[:, :ni] = lag_r, lag_s
        dpsi_r[:, ni:2 * ni], dpsi_s[:, ni:2 * ni] = lag_r, lag_s

        v1inv = jb.inverse_with_check(jb.vandermonde_1d(self.P, self._edge_param))

        def lag1d(t):
            return jb.vandermonde_1d(self.P, t) @ v1inv, jb.grad_vandermonde_1d(self.P, t) @ v1inv

        e1, de1 = lag1d(r)       # bottom edge functions are parameterised by r
        e2, de2 = lag1d(s)       # hypotenuse by s
        e3, de3 = lag1d(-s)      # left edge by -s
        o = 2 * ni
        psi[:, o:o + ne] = e1
        dpsi_r[:, o:o + ne] = de1
        psi[:, o + ne:o + 2 * ne] = e2
        dpsi_s[:, o + ne:o + 2 * ne] = de2
        psi[:, o + 2 * ne:] = e3
        dpsi_s[:, o + 2 * ne:] = -de3

        # Base vector fields e_j(r,s) and their divergence, per column j.
        ex = np.zeros((np_, np_))
        ey = np.zeros((np_, np_))
        dive = np.zeros((np_, np_))
        ex[:, :ni], ey[:, :ni] = (sp * rp)[:, None], (sp * sm)[:, None]
        dive[:, :ni] = ((3.0 * s + 1.0) / 4.0)[:, None]
        ex[:, ni:o], ey[:, ni:o] = (rp * rm)[:, None], (rp * sp)[:, None]
        dive[:, ni:o] = ((3.0 * r + 1.0) / 4.0)[:, None]
        ex[:, o:o + ne], ey[:, o:o + ne] = rp[:, None], sm[:, None]
        dive[:, o:o + ne] = 1.0
        ex[:, o + ne:o + 2 * ne], ey[:, o + ne:o + 2 * ne] = (sr2 * rp)[:, None], (sr2 * sp)[:, None]
        dive[:, o + ne:o + 2 * ne] = sr2
        ex[:, o + 2 * ne:], ey[:, o + 2 * ne:] = rm[:, None], sp[:, None]
        dive[:, o + 2 * ne:] = 1.0

        # V_ij = (psi_j e_j)(r_i, s_i) . dof_i ;  div(psi e) = psi div e + e . grad psi
        self.V = psi * (ex * dof[:, 0:1] + ey * dof[:, 1:2])
        self.VInv = jb.inverse_with_check(self.V)
        div_basis = psi * dive + ex * dpsi_r + ey * dpsi_s
        self.Div = div_basis @ self.VInv
        self.DivInt = self.Div[:ni, :].copy()
