"""The reference's own tests of the 2D modal basis (DG2D/basis_polynomials_test.go) run against the host mirror that
produces V / Vinv / FluxEdgeInterp / the shock finder's operators -- every operator array that crosses the C ABI is built
from this basis, so these pin the INPUTS of both the oracle and the device."""
import numpy as np
import pytest

from gocfd_b200.host.dg2d import jacobi as jb
from gocfd_b200.host.dg2d.elements import JacobiBasis2D, wsj_points

TOL = 1e-6      # the reference's tolerance


def _poly(r, s, p):
    """PolyScalarField.P / .Gradient (DG2D/test_functions.go:98-121): (10 r + s + 10)^p."""
    base = 10.0 * r + s + 10.0
    if p == 0:
        return np.ones_like(r), np.zeros_like(r), np.zeros_like(r)
    return base ** p, 10.0 * p * base ** (p - 1), p * base ** (p - 1)


def test_individual_terms_build_the_vandermonde_matrix():
    """TestJacobiBasis2D_IndividualTerms (basis_polynomials_test.go:31-56): V[i][j] = PolynomialTerm(r_i, s_i, Order2DAtJ[j])
    at P = 2 on the Williams-Shunn-Jameson nodes."""
    r, s, _ = wsj_points(2)
    b = JacobiBasis2D(2, r, s)
    a = np.empty((b.Np, b.Np))
    for j, (i0, j0) in enumerate(b.Order2DAtJ):
        a[:, j] = jb.simplex_2d_p(r, s, i0, j0)
    np.testing.assert_allclose(b.V, a, atol=TOL)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_modal_gradient_of_polynomial_fields(p):
    """TestJacobiBasis2D_Gradient (:58-108): Vr.Vinv and Vs.Vinv differentiate (10r+s+10)^q exactly for q <= P.
    The reference asserts 1e-6 absolute on values up to 21^4; here relative to the field's scale."""
    r, s, _ = wsj_points(p)
    b = JacobiBasis2D(p, r, s)
    for q in range(p + 1):
        f, fr, fs = _poly(r, s, q)
        scale = max(1.0, np.abs(fr).max())
        np.testing.assert_allclose(b.Vr @ b.Vinv @ f, fr, atol=TOL * scale)
        np.testing.assert_allclose(b.Vs @ b.Vinv @ f, fs, atol=TOL * scale)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_nodal_polynomials_are_cardinal_on_their_nodes(p):
    """TestJacobiBasis2D_GetOrthogonalPolynomialAtJ (:110-129): psi_j(r_i, s_i) = delta_ij for P = 1..6."""
    r, s, _ = wsj_points(p)
    b = JacobiBasis2D(p, r, s)
    np.testing.assert_allclose(b.lagrange(r, s), np.eye(b.Np), atol=TOL)


@pytest.mark.parametrize("p", [0, 1, 2, 3, 4, 5, 6])
def test_1d_nodal_basis_and_its_derivative(p):
    """TestJacobiBasis1D_GetOrthogonalPolynomialAtJ / TestLagrangePoly1D (:131-196) on the Legendre zeros: cardinal at the
    nodes, and the derivative matrix differentiates polynomials up to degree P exactly."""
    r = np.polynomial.legendre.leggauss(p + 1)[0]
    v = jb.vandermonde_1d(p, r)
    vr = jb.grad_vandermonde_1d(p, r)
    vinv = np.linalg.inv(v)
    np.testing.assert_allclose(v @ vinv, np.eye(p + 1), atol=TOL)
    dr = vr @ vinv
    for q in range(p + 1):
        f = (0.5 * r + 1.0) ** q
        df = 0.5 * q * (0.5 * r + 1.0) ** (q - 1) if q else np.zeros_like(r)
        np.testing.assert_allclose(dr @ f, df, atol=TOL)
