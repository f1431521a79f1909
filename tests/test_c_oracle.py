"""The C/OpenMP restatement (oracle/c) against the numpy oracle: two independently written
restatements of the reference's stage (inviscid and PerssonC0 dissipation paths) must agree to round-off
(CPU, no GPU)."""
import numpy as np
import pytest

from conftest import mesh_path
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh
from oracle.c_oracle import COracleSolver
from oracle.euler2d_oracle import OracleSolver

TOL = 1e-11      # the same bar as the device path (cancellation in Div·F amplifies round-off order)


def rel_l2(a, b):
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def make(ip_kw, mesh):
    base = dict(CFL=1.0, FluxType="Roe", InitType="Freestream", Minf=0.8, Gamma=1.4, Alpha=1.25,
                FinalTime=100.0, MaxIterations=1000)
    base.update(ip_kw)
    return Euler(InputParameters2D(**base), mesh)


def pair(c):
    a, b = COracleSolver(c.problem), OracleSolver(c.problem)
    a.set_state(c.Q)
    b.set_state(c.Q)
    return a, b


@pytest.mark.parametrize("n", range(5))
@pytest.mark.parametrize("flux", ["average", "lax", "roe", "roe-er"])
def test_rhs_far_field(n, flux):
    c = make(dict(PolynomialOrder=n, FluxType=flux, Minf=0.5), mesh_path("vortex-new.su2"))
    rng = np.random.default_rng(n)
    c.Q = c.Q * (1.0 + 0.02 * rng.standard_normal(c.Q.shape))
    a, b = pair(c)
    assert rel_l2(a.rhs(0), b.rhs(0)) < TOL


@pytest.mark.parametrize("n", [0, 2, 4])
def test_vortex_steps_global_dt(n):
    c = make(dict(PolynomialOrder=n, InitType="IVortex", FinalTime=50.0), structured_tri_mesh(16, 12))
    a, b = pair(c)
    ia, ib = a.step(5), b.step(5)
    assert ia["steps"] == ib["steps"] == 5
    assert abs(ia["time"] - ib["time"]) <= 1e-14 * abs(ib["time"])
    assert rel_l2(a.get_state(), b.get_state()) < TOL
    np.testing.assert_allclose(a.residual(), b.residual(), rtol=1e-10, atol=1e-14)


def test_naca_local_dt_wall_far():
    c = make(dict(PolynomialOrder=1, CFL=2.0, LocalTimeStepping=True, MaxIterations=10), mesh_path("mesh_NACA0012_inv.su2"))
    a, b = pair(c)
    ia, ib = a.step(10), b.step(10)
    assert ia["finished"] and ib["finished"]
    assert rel_l2(a.get_state(), b.get_state()) < TOL


def test_shocktube_in_out_wall_final_time():
    c = make(dict(PolynomialOrder=2, InitType="ShockTube", FluxType="Lax", FinalTime=0.002), mesh_path("sod-aligned-100pts.su2"))
    a, b = pair(c)
    ia, ib = a.step(100), b.step(100)
    assert ia["finished"] and ib["finished"] and ia["steps"] == ib["steps"]
    assert ia["time"] == pytest.approx(0.002, abs=1e-15)
    assert rel_l2(a.get_state(), b.get_state()) < TOL


# ---- PerssonC0 path: the C side follows the Go data flow (materialised DXMetric, DOFX/DOFY,
# Epsilon, DissDOF ...), the numpy side is vectorised differently; agreement is to round-off of the operators.

# N=1: the RT2 divergence operator has entries up to 1.8e3 (its edge points -0.028, 0, 0.028 nearly coincide), so any
# two float64 evaluation orders of Div . DOF differ by ~3e-11 of |RHS| (tests/test_noise_floor.py measures it against
# long double); from N=2 on the floor is ~1e-14.
def _tol(n):
    return 2e-10 if n == 1 else TOL


def _sod(n, **kw):
    base = dict(PolynomialOrder=n, InitType="shocktube", CFL=1.0, FinalTime=0.2, Limiter="persson c0", Kappa=5.0)
    base.update(kw)
    return make(base, mesh_path("sod-aligned-100pts.su2"))


def _smeared(c, width):
    x, _ = c.DFR.solution_xy()
    w = 0.5 * (1.0 - np.tanh((x - 0.503) / width))
    q = np.empty_like(c.Q)
    for n in range(4):
        q[n] = c.FSOut.Qinf[n] + (c.FSIn.Qinf[n] - c.FSOut.Qinf[n]) * w
    return q


@pytest.mark.parametrize("n", [1, 2, 3, 4])
@pytest.mark.parametrize("rk", [0, 2])
def test_dissipation_rhs(n, rk):
    c = _sod(n)
    assert c.problem.Dissipation
    q = _smeared(c, 0.004 if n == 1 else 0.002)
    a, b = pair(c)
    a.set_register(rk, q)                    # rk = 2: the stage input is limited in place (euler.go:605-609)
    b.Q[rk][...] = q
    ra, rb = a.rhs(rk), b.rhs(rk)
    assert b.SigmaScalar.max() > 0.05
    assert rel_l2(ra, rb) < _tol(n)


@pytest.mark.parametrize("n", [2, 4])
def test_sod_steps_with_dissipation(n):
    c = _sod(n, CFL=2.0)
    c.Q = _smeared(c, 0.002)
    a, b = pair(c)
    ia, ib = a.step(10), b.step(10)
    assert ia["steps"] == ib["steps"] == 10
    assert abs(ia["time"] - ib["time"]) <= 1e-12 * abs(ib["time"])
    assert rel_l2(a.get_state(), b.get_state()) < TOL
    np.testing.assert_allclose(a.residual(), b.residual(), rtol=1e-8, atol=1e-12)


def test_naca_transonic_local_dt_with_dissipation():
    """DTVisc carry-over across stages and steps (euler.go:989-999) on the unstructured NACA mesh."""
    c = make(dict(PolynomialOrder=2, CFL=1.0, LocalTimeStepping=True, MaxIterations=12, Minf=0.8, Alpha=2.0,
                  Limiter="PerssonC0", Kappa=4.5), mesh_path("mesh_NACA0012_inv.su2"))
    assert c.problem.Dissipation
    a, b = pair(c)
    a.step(12), b.step(12)
    assert rel_l2(a.get_state(), b.get_state()) < TOL


def _naca_front_case():
    """NACA0012 mesh, N=2, M=0.8, local time stepping, PerssonC0, with a smeared density/energy front at x = 0.6 so that
    the sensor fires in ~400 elements during the first steps (the free-stream start needs thousands of iterations to
    grow a shock): local dt with an ACTIVE viscous limit and the DTVisc carry-over (euler.go:989-999)."""
    c = make(dict(PolynomialOrder=2, CFL=1.0, LocalTimeStepping=True, MaxIterations=100000, Minf=0.8, Alpha=2.0,
                  Limiter="PerssonC0", Kappa=4.5), mesh_path("mesh_NACA0012_inv.su2"))
    x, _ = c.DFR.solution_xy()
    fac = 1.0 + 0.5 * (0.5 * (1.0 - np.tanh((x - 0.6) / 0.004)))
    c.Q[0] *= fac
    c.Q[3] *= fac
    return c


def test_naca_front_local_dt_active_dissipation():
    c = _naca_front_case()
    a, b = pair(c)
    b.rhs(0)
    assert b.SigmaScalar.max() > 0.1 and (b.SigmaScalar > 0.05).sum() > 100
    a.step(2), b.step(2)
    assert (b.DTVisc > 1e-9).sum() > 100           # the viscous branch of CalculateLocalDT was taken
    assert rel_l2(a.get_state(), b.get_state()) < TOL
    np.testing.assert_allclose(a.residual(), b.residual(), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("flux", ["Roe", "Lax", "roe-er", "average"])
@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_dissipation_matrix_fluxes_orders_dt_modes(n, flux, local):
    """Every flux x order x {global, local} dt with an active sensor on the Sod mesh (In / Out / Wall boundaries), 2 RK
    steps: the two restatements were written independently (numpy, vectorised; C, following the Go data flow) and must
    agree to round-off -- measured worst case 6e-14 (N=1), 6e-16 otherwise."""
    c = _sod(n, FluxType=flux, LocalTimeStepping=local, MaxIterations=1000)
    q = _smeared(c, 0.004 if n == 1 else 0.002)
    a, b = pair(c)
    a.set_state(q), b.set_state(q)
    a.step(2), b.step(2)
    assert np.isfinite(b.get_state()).all()
    assert rel_l2(a.get_state(), b.get_state()) < (1e-11 if n == 1 else 1e-13)
