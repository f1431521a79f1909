"""The C/OpenMP restatement (oracle/c) against the numpy oracle: two independently written
restatements of the reference's inviscid stage must agree to round-off (CPU, no GPU)."""
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


def test_refuses_dissipation():
    c = make(dict(PolynomialOrder=2, InitType="ShockTube", Limiter="PerssonC0", Kappa=3.0), mesh_path("sod-aligned-100pts.su2"))
    with pytest.raises(ValueError):
        COracleSolver(c.problem)
