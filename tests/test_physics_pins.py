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
  * the five-stage combination of rkAdvance (euler.go:511-563) is fourth-order accurate on a nonlinear ODE and has the
    stability polynomial of the Spiteri-Ruuth SSPRK(5,4) scheme;
  * the time step is the CFL condition: CFL x hK / (|V| + c) globally and per element on a uniform stream;
  * the modal sensor: exactly Persson & Peraire's energy ratio at N = 1; blind to everything below the top modes at any N;
  * WallBC (bcs.go:11-23): a gas at rest inside solid walls stays at rest, and a centred pressure pulse in a walled box
    stays mirror symmetric;
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


# ---- SSP-RK(5,4): the combination rkAdvance applies (euler.go:511-563) is a fourth-order scheme ------------------------
class _ScalarODE:
    """The numpy oracle's rk_advance with the three calls that produce RHSQ replaced by an ODE right-hand side, so that the
    five-stage combination -- registers, coefficients, the Residual carried from stage 4 into stage 5 -- is exercised
    exactly as StepWorker drives it, on a problem with a closed-form solution."""

    def __new__(cls, f, y0, dt):
        from oracle.euler2d_oracle import OracleSolver

        class Solver(OracleSolver):
            def set_rt_flux_internal(self, q):
                self._stage_input = q

            def set_rt_flux_on_edges(self):
                pass

            def rhs_internal_points(self):
                self.RHSQ = f(np.asarray(self._stage_input))

        c = Euler(InputParameters2D(CFL=1.0, FluxType="Lax", InitType="Freestream", PolynomialOrder=0, FinalTime=1.0,
                                    MaxIterations=10, Gamma=1.4, Minf=0.5), mesh_path("test_tris_two.neu"))
        s = Solver(c.problem)
        s.set_state(np.full_like(c.Q, y0))
        s.DT[...] = dt
        return s


def _integrate(f, y0, t_end, steps):
    s = _ScalarODE(f, y0, t_end / steps)
    for _ in range(steps):
        for rk in range(5):
            s.rk_advance(rk)
    return float(s.get_state()[0, 0, 0])


def test_ssp_rk54_combination_is_fourth_order_on_a_nonlinear_ode():
    """y' = -y^2, y(0) = 1, y(t) = 1 / (1 + t): a nonlinear problem, so all eight order conditions up to order four are
    in play.  Halving dt divides the error by 2^4 (observed order within 3.9..4.1), which neither a third-order nor a
    mis-typed coefficient survives."""
    exact = 1.0 / 3.0
    errs = [abs(_integrate(lambda y: -y * y, 1.0, 2.0, n) - exact) for n in (20, 40, 80)]
    orders = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(3.9 < o < 4.1 for o in orders), (errs, orders)
    assert errs[-1] < 2e-9


def test_ssp_rk54_stability_polynomial_is_spiteri_ruuth():
    """On y' = lambda y one step multiplies y by the stability polynomial R(z), z = lambda dt.  For the Spiteri-Ruuth
    SSPRK(5,4) scheme R(z) = 1 + z + z^2/2 + z^3/6 + z^4/24 + 0.0044777 z^5: the first five coefficients are the order
    conditions; the z^5 coefficient (= the product of the five stage weights, 1/223.3) identifies the scheme among fourth-order five-stage methods."""
    zs = np.array([0.05, 0.1, 0.2, 0.4, 0.8, -0.3])
    r = np.array([_integrate(lambda y, z=z: z * y, 1.0, 1.0, 1) for z in zs])
    taylor4 = 1 + zs + zs ** 2 / 2 + zs ** 3 / 6 + zs ** 4 / 24
    c5 = (r - taylor4) / zs ** 5
    np.testing.assert_allclose(c5, c5[0], rtol=1e-6)          # a pure z^5 remainder: R is a degree-5 polynomial
    assert abs(c5[0] - 0.391752226571890 * 0.368410593050371 * 0.251891774271694 * 0.544974750228521 * 0.226007483236906) < 1e-9


# ---- WallBC (bcs.go:11-23): only pressure acts on a wall ---------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 3])
def test_gas_at_rest_in_a_walled_box_stays_at_rest(n):
    """A uniform gas at rest inside solid walls: the wall flux is (0, p nx, p ny, 0), which is exactly the physical flux of
    the interior state, so the divergence vanishes identically -- for any correct WallBC, edge orientation bookkeeping and
    RT divergence, and for no sign error in any of them."""
    from oracle.euler2d_oracle import OracleSolver
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=n, FinalTime=1.0, MaxIterations=10,
                                Gamma=1.4, Minf=0.0), structured_tri_mesh(5, 4, tag="wall"))
    o = OracleSolver(c.problem)
    o.set_state(c.Q)
    assert np.abs(c.Q[1]).max() == 0.0 and np.abs(c.Q[2]).max() == 0.0
    assert np.abs(o.rhs(0)).max() < 1e-12
    o.step(2)
    np.testing.assert_allclose(o.get_state(), c.Q, rtol=0, atol=2e-12)


def test_wall_reflects_a_normal_pressure_pulse_symmetrically():
    """Mirror symmetry: a pressure pulse centred on the vertical mid-line of a walled box must stay mirror symmetric --
    density and energy even, x-momentum odd -- which ties the wall treatment on the left and right walls (opposite normals,
    opposite edge orientations in the owner's traversal) to each other."""
    nx_, ny_ = 8, 4
    c = Euler(InputParameters2D(CFL=0.5, FluxType="Lax", InitType="Freestream", PolynomialOrder=2, FinalTime=10.0, MaxIterations=100,
                                Gamma=1.4, Minf=0.0), structured_tri_mesh(nx_, ny_, -1.0, 1.0, -0.5, 0.5, tag="wall"))
    x, y = c.DFR.solution_xy()
    bump = 0.2 * np.exp(-20.0 * x * x)
    q = c.Q.copy()
    q[0] = q[0] * (1.0 + bump)
    q[3] = q[3] * (1.0 + 1.4 * bump)
    o = COracleSolver(c.problem)
    o.set_state(q)
    o.step(12)
    out = o.get_state()
    o.close()
    # the triangulation itself is not mirror symmetric (all diagonals run the same way), so compare moments, not nodes
    w = c.problem.Jdet[None, :] * np.ones_like(x)
    left, right = x < 0, x > 0
    mass_l, mass_r = (out[0] * w)[left].sum(), (out[0] * w)[right].sum()
    mom_l, mom_r = (out[1] * w)[left].sum(), (out[1] * w)[right].sum()
    assert abs(mass_l - mass_r) < 2e-3 * abs(mass_l)
    assert abs(mom_l + mom_r) < 2e-2 * max(abs(mom_l), abs(mom_r)) and abs(mom_l) > 1e-4
    assert mom_l < 0 < mom_r                                   # the pulse pushes gas outwards on both sides


# ---- the time step is a CFL number (euler.go:945-1002, edges.go:246-289) ---------------------------------------------
def test_global_and_local_dt_are_the_cfl_condition_on_a_uniform_stream():
    """With a uniform stream the wave speed |V| + c is the same number everywhere, so the time step must be
    CFL x (length scale) / (|V| + c) with the length scale hK = EdgeLenMax / (N+1)^2 (the reference's choice,
    DG2D/dfr_startup.go:125-147): globally the smallest hK of the mesh, locally the smallest hK among an element and the
    owners of its three edges."""
    from oracle.euler2d_oracle import OracleSolver
    for local, cfl, n in ((False, 0.7, 2), (True, 1.3, 1)):
        c = Euler(InputParameters2D(CFL=cfl, FluxType="Roe", InitType="Freestream", PolynomialOrder=n, FinalTime=100.0,
                                    MaxIterations=10, Gamma=1.4, Minf=0.5, Alpha=3.0, LocalTimeStepping=local),
                  mesh_path("mesh_NACA0012_inv.su2"))
        p = c.problem
        o = OracleSolver(p)
        o.set_state(c.Q)
        rho, ru, rv, e = (c.Q[v][0, 0] for v in range(4))
        pr = (G - 1) * (e - 0.5 * (ru * ru + rv * rv) / rho)
        wave = np.hypot(ru, rv) / rho + np.sqrt(G * pr / rho)
        hk = p.EdgeLenMax / float((n + 1) ** 2)
        o.stage(0)                                          # stage 1 computes dt from the (uniform) input state
        if not local:
            assert o.GlobalDT == pytest.approx(cfl * hk.min() / wave, rel=1e-12)
        else:
            h_adj = np.minimum.reduce([hk[p.edge_kL[p.EtoEdge[:, e]]] for e in range(3)])
            np.testing.assert_allclose(o.DT, cfl * h_adj / wave, rtol=1e-12)
            assert o.DT.max() / o.DT.min() > 50             # a real spread of cell sizes (airfoil surface vs far field)


# ---- the modal sensor (UpdateSeMoment, dissipation.go:455-488; ModeAliasShockFinder, DG2D/dfr_shock_capturing.go:70-105) --
def _sensor_case(n):
    from oracle.euler2d_oracle import OracleSolver
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n, FinalTime=0.2, MaxIterations=10,
                                Gamma=1.4, Limiter="persson c0", Kappa=5.0), mesh_path("sod-aligned-100pts.su2"))
    return c, OracleSolver(c.problem)


def test_sensor_is_the_persson_peraire_energy_ratio_at_n1():
    """At N = 1 the WSJ cubature integrates the products of basis functions exactly, the reference's `MassMatrix` is the
    identity in modal space, and Se is exactly log10(energy in the linear modes / total modal energy) -- the definition of
    Persson & Peraire's indicator, evaluated here from the modal coefficients alone."""
    c, o = _sensor_case(1)
    p = c.problem
    rho = 1.0 + 0.3 * np.random.default_rng(1).random((p.NpInt, p.K))
    o.update_se_moment(rho)
    uh = p.Vinv @ rho
    top = np.array(c.DFR.SolutionElement.JB2D.OrderAtJ) == 1
    np.testing.assert_allclose(o.Se, np.log10((uh[top] ** 2).sum(axis=0) / (uh ** 2).sum(axis=0)), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_sensor_sees_only_the_top_modes(n):
    """For any order: a density that is a polynomial of degree N-1 inside every element has no top-mode content, so the
    numerator vanishes (Se at round-off level, far below the ramp: sigma = 0); adding a top-order mode of relative size
    0.1 lifts Se to about log10(0.01) = -2, inside the ramp [S0 - Kappa, S0 + Kappa] of the limiter."""
    c, o = _sensor_case(n)
    p = c.problem
    orders = np.array(c.DFR.SolutionElement.JB2D.OrderAtJ)
    rng = np.random.default_rng(n)
    uh = np.zeros((p.NpInt, p.K))
    uh[0] = 3.0
    uh[(orders > 0) & (orders < n)] = 0.2 * rng.standard_normal(((orders > 0) & (orders < n)).sum())[:, None]
    with np.errstate(all="ignore"):
        o.update_se_moment(p.V @ uh)
        o.update_shock_finder_sigma()
    assert np.nan_to_num(o.Se, nan=-99.0, neginf=-99.0).max() < -20 and o.SigmaScalar.max() == 0.0
    uh[orders == n] = 0.3 / np.sqrt((orders == n).sum())
    o.update_se_moment(p.V @ uh)
    o.update_shock_finder_sigma()
    assert np.all((o.Se > -2.6) & (o.Se < -1.4))
    assert np.all((o.SigmaScalar > 0.0) & (o.SigmaScalar < 1.0))
