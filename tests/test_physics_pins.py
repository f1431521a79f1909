"""Pins for the parts of the path that no reference test covers (Roe / Lax flux values, RiemannBC,
the SSP coefficients in action, dt logic, after-N-steps fields).  The Go solver cannot be run here, so these are pinned
by what they must equal mathematically or physically -- independent of the Go source and of our reading of it:

  * RoeFlux (fluxes.go:284-413) == the textbook Roe flux (Roe average, four wave strengths, no entropy fix);
  * LaxFlux (fluxes.go:161-190) == central flux + 1/2 max(|V|+c) (qL - qR);
  * every numerical flux is consistent, F(q, q, n) = f(q).n (incl. Roe-ER, which is otherwise reproduced as written,
    fluxes.go:415-503), and Roe / Lax / average are conservative;
  * RiemannBC (bcs.go:70-133) leaves the free stream fixed; its supersonic branch copies the exterior / interior state;
  * the inviscid stage + SSP-RK(5,4) + global dt + analytic-vortex boundary converge to the exact isentropic vortex at
    high order (error ratio between two meshes), for N = 1..4;
  * the PerssonC0 path carries a Sod shock to t = 0.1 and lands on the exact Riemann solution (centre-line samples of
    OutputFinal against SOD_Exact).
CPU only; the time-dependent cases run through the C restatement (oracle/c), which the numpy oracle and the CUDA path are
compared with at 1e-11 elsewhere.
"""
import numpy as np
import pytest

from conftest import mesh_path
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh
from host_standin.sod_shock_tube import SODExact, SODShockTube
from oracle.c_oracle import COracleSolver
from oracle.euler2d_oracle import avg_flux, flux_calc_base, lax_flux, riemann_bc, roe_er_flux, roe_flux

G = 1.4


def _phys(q):
    r, ru, rv, e = q
    u, v = ru / r, rv / r
    p = (G - 1) * (e - 0.5 * r * (u * u + v * v))
    return np.array([ru, ru * u + p, ru * v, u * (e + p)]), np.array([rv, rv * u, rv * v + p, v * (e + p)]), p


def _roe_textbook(ql, qr, nx, ny):
    fl, gl, pl = _phys(ql)
    fr, gr, pr = _phys(qr)
    rl, rr = ql[0], qr[0]
    ul, vl, ur, vr = ql[1] / rl, ql[2] / rl, qr[1] / rr, qr[2] / rr
    hl, hr = (ql[3] + pl) / rl, (qr[3] + pr) / rr
    sl, sr = np.sqrt(rl), np.sqrt(rr)
    u, v, h = (sl * ul + sr * ur) / (sl + sr), (sl * vl + sr * vr) / (sl + sr), (sl * hl + sr * hr) / (sl + sr)
    a = np.sqrt((G - 1) * (h - 0.5 * (u * u + v * v)))
    rho = sl * sr
    un = u * nx + v * ny
    dun = (ur - ul) * nx + (vr - vl) * ny
    dut = -(ur - ul) * ny + (vr - vl) * nx
    dr, dp = rr - rl, pr - pl
    a1, a2, a3, a4 = (dp - rho * a * dun) / (2 * a * a), dr - dp / (a * a), rho * dut, (dp + rho * a * dun) / (2 * a * a)
    k1 = np.array([1, u - a * nx, v - a * ny, h - a * un])
    k2 = np.array([1, u, v, 0.5 * (u * u + v * v)])
    k3 = np.array([0, -ny, nx, -u * ny + v * nx])
    k4 = np.array([1, u + a * nx, v + a * ny, h + a * un])
    d = abs(un - a) * a1 * k1 + abs(un) * a2 * k2 + abs(un) * a3 * k3 + abs(un + a) * a4 * k4
    return 0.5 * ((fl + fr) * nx + (gl + gr) * ny) - 0.5 * d


def _states(seed, jump=0.1, count=200):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        ql = np.array([1 + 0.3 * rng.random(), 0.5 * rng.standard_normal(), 0.5 * rng.standard_normal(), 2.5 + rng.random()])
        qr = ql * (1 + jump * rng.standard_normal(4))
        th = rng.random() * 2 * np.pi
        yield ql, qr, np.cos(th), np.sin(th)


def _call(f, ql, qr, nx, ny):
    return np.array(f(G, [np.array([x]) for x in ql], [np.array([x]) for x in qr], np.array([nx]), np.array([ny]))).ravel()


def test_roe_flux_is_the_textbook_roe_flux():
    for ql, qr, nx, ny in _states(0):
        want = _roe_textbook(ql, qr, nx, ny)
        assert np.abs(_call(roe_flux, ql, qr, nx, ny) - want).max() <= 1e-13 * np.abs(want).max()


def test_lax_flux_is_central_plus_max_wave_speed_jump():
    for ql, qr, nx, ny in _states(1):
        fl, gl, pl = _phys(ql)
        fr, gr, pr = _phys(qr)
        cl, cr = np.sqrt(G * pl / ql[0]), np.sqrt(G * pr / qr[0])
        vl, vr = np.hypot(ql[1], ql[2]) / ql[0], np.hypot(qr[1], qr[2]) / qr[0]
        want = 0.5 * ((fl + fr) * nx + (gl + gr) * ny) + 0.5 * max(vl + cl, vr + cr) * (ql - qr)
        assert np.abs(_call(lax_flux, ql, qr, nx, ny) - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("flux", [roe_flux, lax_flux, avg_flux, roe_er_flux])
def test_numerical_fluxes_are_consistent(flux):
    """F(q, q, n) = f(q) . n."""
    for ql, _, nx, ny in _states(2, count=50):
        f, g, _ = _phys(ql)
        want = f * nx + g * ny
        assert np.abs(_call(flux, ql, ql, nx, ny) - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("flux", [roe_flux, lax_flux, avg_flux])
def test_numerical_fluxes_are_conservative(flux):
    """F(qL, qR, n) = -F(qR, qL, -n): what one element loses through an edge the neighbour gains."""
    for ql, qr, nx, ny in _states(3, count=50):
        a, b = _call(flux, ql, qr, nx, ny), _call(flux, qr, ql, -nx, -ny)
        assert np.abs(a + b).max() <= 1e-13 * np.abs(a).max()


def test_riemann_bc_fixed_point_and_supersonic_branches():
    from gocfd_b200.host.euler2d import FreeStream
    fs = FreeStream.from_mach(0.5, G, 3.0)
    tup = tuple(fs.as_array())
    q = [np.array([x]) for x in fs.Qinf]
    for th in np.linspace(0, 2 * np.pi, 13):
        out = riemann_bc(tup, q, fs.Qinf, np.array([np.cos(th)]), np.array([np.sin(th)]))
        np.testing.assert_allclose(np.array(out).ravel(), fs.Qinf, rtol=1e-13, atol=1e-14)
    # supersonic free stream: inflow copies the exterior state, outflow keeps the interior one (bcs.go:122-131)
    fs2 = FreeStream.from_mach(2.0, G, 0.0)
    qi = [np.array([1.1]), np.array([2.3]), np.array([0.1]), np.array([4.5])]
    inflow = riemann_bc(tuple(fs2.as_array()), qi, fs2.Qinf, np.array([-1.0]), np.array([0.0]))
    outflow = riemann_bc(tuple(fs2.as_array()), qi, fs2.Qinf, np.array([1.0]), np.array([0.0]))
    np.testing.assert_allclose(np.array(inflow).ravel(), fs2.Qinf)
    np.testing.assert_allclose(np.array(outflow).ravel(), [1.1, 2.3, 0.1, 4.5])


def _vortex_error(n, nx, flux, cfl, t_final=0.5):
    ip = InputParameters2D(CFL=cfl, FluxType=flux, InitType="IVortex", PolynomialOrder=n, FinalTime=t_final,
                           MaxIterations=10 ** 6, Gamma=G, Minf=0.1)
    c = Euler(ip, structured_tri_mesh(nx, nx, tag="wall"))
    o = COracleSolver(c.problem)
    o.set_state(c.Q)
    info = o.step(10 ** 6)
    assert info["finished"] and info["time"] == pytest.approx(t_final, abs=1e-12)
    q = o.get_state()
    o.close()
    x, y = c.DFR.solution_xy()
    exact = np.zeros_like(q)
    for i in range(x.shape[0]):
        for k in range(x.shape[1]):
            exact[:, i, k] = c.AnalyticSolution.get_state_c(info["time"], x[i, k], y[i, k])
    w = c.DFR.Jdet[None, None, :]
    return float(np.sqrt((w * (q - exact) ** 2).sum() / (w * np.ones_like(q)).sum()))


@pytest.mark.parametrize("n,flux,coarse,fine,cfl,min_order,max_fine_error", [
    (1, "Roe", 24, 48, 0.5, 1.6, 5e-3),
    (2, "Roe", 24, 48, 0.5, 2.5, 8e-4),
    (1, "Lax", 24, 48, 0.5, 1.8, 5e-3),
    (2, "Lax", 24, 48, 0.5, 2.2, 8e-4),
    (3, "Lax", 16, 32, 0.25, 3.2, 6e-4),
    (4, "Lax", 12, 24, 0.25, 3.7, 6e-4),
])
def test_isentropic_vortex_converges_at_high_order(n, flux, coarse, fine, cfl, min_order, max_fine_error):
    """Measured orders (density-weighted L2 of all four variables at t = 0.5): Roe 1.80 / 2.75, Lax 2.13 / 2.46 / 3.55 /
    4.08 for N = 1..4 -- a wrong SSP coefficient, flux sign, edge orientation, dt or boundary state destroys them."""
    e1, e2 = _vortex_error(n, coarse, flux, cfl), _vortex_error(n, fine, flux, cfl)
    assert e2 < max_fine_error
    assert np.log2(e1 / e2) > min_order


def test_sod_with_persson_dissipation_lands_on_the_exact_solution():
    """C3: shipped Sod mesh, N=2, PerssonC0, to t = 0.1; centre-line samples (SODShockTube, as OutputFinal takes them)
    against SOD_Exact (measured: plateaus 0.421 / 0.2655 for 0.4263 / 0.2656, mean |error| 0.0075, shock at 0.68 for
    0.675): plateau values within 2 %, mean density error below 1.5 % of the jump, shock position within 2 cells."""
    ip = InputParameters2D(CFL=2.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=2, FinalTime=0.1,
                           MaxIterations=10 ** 6, Gamma=G, Limiter="persson c0", Kappa=5.0)
    c = Euler(ip, mesh_path("sod-aligned-100pts.su2"))
    o = COracleSolver(c.problem)
    o.set_state(c.Q)
    info = o.step(10 ** 6)
    assert info["finished"]
    q = o.get_state()
    o.close()
    assert np.isfinite(q).all() and q[0].min() > 0.1
    st = SODShockTube(4 * c.DFR.K // 5, c.DFR)
    st.interpolate_fields(q)
    sod = SODExact(info["time"])
    exact = np.array([sod.getx(x)[0] for x in st.XLocations])
    assert np.abs(st.Rho - exact).mean() < 0.015 * (1.0 - 0.125)
    x = st.XLocations
    mid = (x > sod.x2 + 0.02) & (x < sod.x3 - 0.04)          # between rarefaction tail and (smeared) contact
    post = (x > sod.x3 + 0.03) & (x < sod.x4 - 0.02)         # between contact and shock
    assert mid.sum() > 10 and post.sum() > 5
    assert np.abs(st.Rho[mid] - sod.rho_middle).max() < 0.02 * sod.rho_middle
    assert np.abs(st.Rho[post] - sod.post_s.rho).max() < 0.02 * sod.post_s.rho
    # shock position: where the sampled density crosses the mean of the two states around x4
    level = 0.5 * (sod.post_s.rho + 0.125)
    xs = x[x > sod.x3 + 0.02]
    rs = st.Rho[x > sod.x3 + 0.02]
    x_shock = xs[np.argmax(rs < level)]
    assert abs(x_shock - sod.x4) < 0.02
