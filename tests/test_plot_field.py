"""Field read-back path: Euler.GetPlotField restated in the oracle and on the device.

The reference has no asserting test for GraphInterp / GetPlotField (its tests only plot), so the oracle is pinned
by what the construction guarantees: GraphInterp reproduces every polynomial of degree <= N exactly at the graph
nodes (DG2D/dfr_startup.go:62-63), the three vertex rows are the mean of their boundary neighbours
(DG2D/graphics_support2.go:184-199), and the flow functions are the closed forms of fluids.go:289-336.
"""
import numpy as np
import pytest

from conftest import mesh_path
from gocfd_b200.host.dg2d.dfr2d import DFR2D
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh
from oracle import euler2d_oracle as ora


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_graph_interp_reproduces_polynomials(n):
    d = DFR2D(n)
    gi = d.graph_interp()
    gr, gs = d.graph_rs()
    el = d.SolutionElement
    ne = d.FluxElement.NpEdge
    assert gi.shape == (3 * (1 + ne) + el.Np, el.Np)
    for a in range(n + 1):
        for b in range(n + 1 - a):
            f_nodes = el.R ** a * el.S ** b
            np.testing.assert_allclose(gi @ f_nodes, gr ** a * gs ** b, atol=2e-12)
    # graph node order: vertex, NpEdge edge points (x3), then the interior points themselves
    np.testing.assert_allclose(gi[3 * (1 + ne):], np.eye(el.Np), atol=1e-12)
    assert (gr[0], gs[0]) == (-1.0, -1.0) and (gr[1 + ne], gs[1 + ne]) == (1.0, -1.0)


def _case(n=2, **kw):
    base = dict(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=n, FinalTime=1.0, MaxIterations=10,
                Gamma=1.4, Minf=0.4)
    base.update(kw)
    return Euler(InputParameters2D(**base), structured_tri_mesh(7, 5))


def test_oracle_plot_field_closed_forms_and_vertex_rule():
    c = _case(3)
    p = c.problem
    gi = c.DFR.graph_interp()
    ne = p.NpEdge
    rho, ru, rv, e = c.Q
    fs = tuple(p.FSFar.as_array())
    pres = (p.Gamma - 1.0) * (e - 0.5 * (ru * ru + rv * rv) / rho)
    np.testing.assert_allclose(ora.flow_function(fs, c.Q, ora.FF_StaticPressure), pres, rtol=1e-14)
    np.testing.assert_allclose(ora.flow_function(fs, c.Q, ora.FF_Mach),
                               np.sqrt((ru * ru + rv * rv) / rho ** 2) / np.sqrt(p.Gamma * pres / rho), rtol=1e-13)
    np.testing.assert_allclose(ora.flow_function(fs, c.Q, ora.FF_PressureCoefficient), (pres - fs[5]) / fs[6], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ora.flow_function(fs, c.Q, ora.FF_Entropy), np.log(pres) - p.Gamma * np.log(rho), rtol=1e-12, atol=1e-14)
    out = ora.plot_field(p, c.Q, ora.FF_Density, gi)
    assert out.shape == (p.K, 3 * (1 + ne) + p.NpInt)
    raw = (gi @ rho).T
    npe = ne + 2
    for n_edge in range(3):
        iv = n_edge * (npe - 1)
        ivm = 3 * (npe - 1) - 1 if n_edge == 0 else iv - 1
        np.testing.assert_array_equal(out[:, iv], 0.5 * (raw[:, iv + 1] + raw[:, ivm]))
    keep = np.setdiff1d(np.arange(out.shape[1]), [0, npe - 1, 2 * (npe - 1)])
    np.testing.assert_array_equal(out[:, keep], raw[:, keep])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4])
def test_device_plot_field_matches_oracle(n):
    from gocfd_b200 import lib
    c = _case(n)
    p = c.problem
    gi = c.DFR.graph_interp()
    dev = lib.Dfr2d(p)
    dev.set_state(c.Q)
    dev.step(2)
    q = dev.get_state()
    for ff in range(14):
        want = ora.plot_field(p, q, ff, gi)
        got = dev.plot_field(ff, gi)
        assert got.dtype == np.float32 and got.shape == want.shape
        # float64 arithmetic narrowed to float32 at the end on both sides: at most one float32 ulp apart
        np.testing.assert_allclose(got, want.astype(np.float32), rtol=2.5e-7, atol=1e-7)
    with pytest.raises(lib.Dfr2dError):
        dev.plot_field(99, gi)
    with pytest.raises(lib.Dfr2dError):
        dev.plot_field(0, gi[:-1])
    with pytest.raises(lib.Dfr2dError):
        dev.epsilon_field()                  # no limiter, no epsilon (c.Dissipation == nil)
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 4])
@pytest.mark.parametrize("diss", [False, True])
def test_device_shock_function_and_epsilon_fields(n, diss):
    """The plot fields that do not go through GetFlowFunction (plot.go:30-53): ShockFunction (ShockIndicator of the density,
    Kappa 2 without the limiter, ip.Kappa with it, S0 = 4/N^4) through the GraphInterp path, and -- with the limiter --
    the two epsilon fields [NpFlux x K] as the last stage left them."""
    from conftest import mesh_path
    from gocfd_b200 import lib
    from gocfd_b200.host.euler2d import Euler
    from gocfd_b200.host.input_parameters import InputParameters2D
    from oracle.euler2d_oracle import OracleSolver
    kw = dict(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n, FinalTime=0.2, MaxIterations=100, Gamma=1.4)
    if diss:
        kw.update(Limiter="persson c0", Kappa=5.0)
    c = Euler(InputParameters2D(**kw), mesh_path("sod-aligned-100pts.su2"))
    p = c.problem
    x, _ = c.DFR.solution_xy()
    w = 0.5 * (1.0 - np.tanh((x - 0.503) / (0.004 if n == 1 else 0.002)))
    q0 = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w for v in range(4)])
    gi = c.DFR.graph_interp()
    dev, o = lib.Dfr2d(p), OracleSolver(p)
    # ShockFunction of a sharp front (no stepping needed: the field is a function of c.Q alone)
    ws = 0.5 * (1.0 - np.tanh((x - 0.503) / 0.0003))
    q = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * ws for v in range(4)])
    dev.set_state(q)
    want = ora.plot_field(p, q, ora.FF_ShockFunction, gi)
    got = dev.plot_field(ora.FF_ShockFunction, gi)
    if n >= 2 and not diss:
        assert want.max() > 0.1 and want.min() == 0.0       # the front is flagged (Kappa 2), the plateaus are not
    # sigma is a function of log10 of a ratio of sums: summation order moves it at the 1e-12 level, far below float32
    np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-6, atol=1e-6)
    dev.set_state(q0)
    o.set_state(q0)
    dev.step(2), o.step(2)
    if diss:
        assert o.EpsilonScalar.max() > 0
        np.testing.assert_allclose(dev.epsilon_field(False), np.tile(o.EpsilonScalar, (p.NpFlux, 1)), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(dev.epsilon_field(True), o.Epsilon, rtol=1e-9, atol=1e-14)
    dev.close()


@pytest.mark.gpu
def test_device_plot_field_partitions_cover_the_mesh():
    from gocfd_b200 import lib
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=2, FinalTime=1.0,
                                MaxIterations=10, Gamma=1.4, Minf=0.8, Alpha=2.0), mesh_path("mesh_NACA0012_inv.su2"))
    p = c.problem
    gi = c.DFR.graph_interp()
    rng = np.random.default_rng(5)
    q = c.Q * (1.0 + 0.05 * rng.standard_normal(c.Q.shape))
    one = lib.Dfr2d(p)
    one.set_state(q)
    ref = one.plot_field(ora.FF_Mach, gi)
    out = np.zeros_like(ref)
    devs = [lib.Dfr2d(p, n_parts=3, part=r) for r in range(3)]
    for d in devs:
        d.set_state(q)
        d.plot_field(ora.FF_Mach, gi, out)
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_allclose(ref, ora.plot_field(p, q, ora.FF_Mach, gi).astype(np.float32), rtol=2.5e-7, atol=1e-7)
    for d in devs + [one]:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case,n", [("IVortex", 2), ("IVortex", 4), ("Freestream", 1), ("shocktube", 3)])
def test_device_initial_condition_matches_host(case, n):
    """InitializeSolution evaluated on the device equals the host mirror's c.Q (which the oracle KATs
    pin: vortex centre state, freestream values) to round-off of exp/pow."""
    from gocfd_b200 import lib
    mesh = mesh_path("sod-aligned-100pts.su2") if case == "shocktube" else structured_tri_mesh(16, 12)
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType=case, PolynomialOrder=n, FinalTime=1.0, MaxIterations=5,
                                Gamma=1.4, Minf=0.3, Alpha=1.0), mesh)
    dev = lib.Dfr2d(c.problem)
    el = c.DFR.SolutionElement
    dev.init_state(c.problem.Case, c.DFR.VX, c.DFR.VY, c.DFR.EToV, el.R, el.S)
    got = dev.get_state()
    np.testing.assert_allclose(got, c.Q, rtol=2e-14, atol=1e-15)
    # and the run that follows is the same run
    a = dev.step(2)
    ref = lib.Dfr2d(c.problem)
    ref.set_state(c.Q)
    b = ref.step(2)
    assert a["steps"] == b["steps"]
    np.testing.assert_allclose(dev.get_state(), ref.get_state(), rtol=1e-12, atol=1e-14)
    dev.close(), ref.close()


# ---- gradient plot fields: XGradient* / YGradient* (plot.go:54-77, fluids.go:227-234) --------------------------------
GRADIENT_FIELDS = [200, 201, 202, 203, 300, 301, 302, 303]


@pytest.mark.parametrize("n", [1, 2, 4])
def test_oracle_gradient_plot_field_is_exact_for_polynomials(n):
    """What the construction guarantees (the reference's TestDissipation2 / TestGradient check the same chain at 1e-6):
    with continuous edge values the RT gradient of a polynomial of degree <= N is its exact derivative at every RT point.
    No limiter here: DXMetric / DYMetric exist for every run (NewDFR2D, DG2D/dfr_startup.go:213-254)."""
    c = _case(n, InitType="Freestream")
    p = c.problem
    x, y = c.DFR.flux_xy()
    xs, ys = x[:p.NpInt], y[:p.NpInt]
    o = ora.OracleSolver(p)
    q = np.stack([1.0 + 0.1 * xs ** m + 0.2 * ys ** m + (0.05 * xs * ys if n >= 2 else 0.0) for m in (1, n, 1, n)])
    o.set_state(q)
    o.interpolate_to_edges(o.Q[0])
    for v in range(4):
        o.EdgeFlux[1][v] = o.Q_Face[v][o.rowsL, o.kLc]
    for v, m in enumerate((1, n, 1, n)):
        cross = 0.05 if n >= 2 else 0.0
        np.testing.assert_allclose(o.gradient_plot_field(200 + v), 0.1 * m * x ** (m - 1) + cross * y, atol=2e-10)
        np.testing.assert_allclose(o.gradient_plot_field(300 + v), 0.2 * m * y ** (m - 1) + cross * x, atol=2e-10)
    with pytest.raises(ValueError):
        o.gradient_plot_field(204)


def test_oracle_gradient_field_uses_the_stale_edge_store():
    """plot.go:69 passes the current c.Q, but the edge rows come from the EdgeQValues store, which the last
    CalculateEdgeEulerFlux filled from the input of stage 5 (edges.go:344-350): before any step the store is zero."""
    c = _case(2)
    o = ora.OracleSolver(c.problem)
    o.set_state(c.Q)
    g0 = o.gradient_plot_field(200)
    un = np.concatenate([c.Q[0], c.Q[0], np.zeros((3 * c.problem.NpEdge, c.problem.K))])
    np.testing.assert_allclose(g0, c.problem.Div @ (o.DXMetric * un), rtol=1e-14, atol=1e-14)
    o.step(1)
    stale = o.EdgeFlux[1][0].copy()
    o.interpolate_to_edges(o.Q[0])
    fresh = o.Q_Face[0][o.rowsL, o.kLc]
    assert np.abs(stale - fresh).max() > 1e-6          # the store is NOT the interpolation of the plotted state


def _gradient_fields(dev, p):
    return {pf: dev.gradient_field(pf) for pf in GRADIENT_FIELDS}


def _check_gradients(got, o, tol=1e-11):
    for pf in GRADIENT_FIELDS:
        want = o.gradient_plot_field(pf)
        assert np.abs(got[pf] - want).max() <= tol * max(1.0, np.abs(want).max()), pf


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4])
def test_device_gradient_fields_match_oracle(n):
    """All eight fields after 0, 1 and 3 steps on the vortex (all-IVortex boundary, local dt at odd orders)."""
    from gocfd_b200 import lib
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=n, FinalTime=50.0,
                                MaxIterations=100, Gamma=1.4, Minf=0.4, LocalTimeStepping=bool(n % 2)), structured_tri_mesh(14, 9))
    p = c.problem
    dev, o = lib.Dfr2d(p), ora.OracleSolver(p)
    dev.set_state(c.Q), o.set_state(c.Q)
    with pytest.raises(lib.Dfr2dError, match="dfr2d_capture_edge_values"):
        dev.gradient_field(200)
    dev.capture_edge_values(True)
    _check_gradients(_gradient_fields(dev, p), o)        # store still zero on both sides
    for steps in (1, 2):
        dev.step(steps), o.step(steps)
        _check_gradients(_gradient_fields(dev, p), o)
    with pytest.raises(lib.Dfr2dError, match="200..203"):
        dev.gradient_field(204)
    dev.close()


@pytest.mark.gpu
def test_device_gradient_fields_capture_window_and_finished_run():
    """Capture switched off keeps what was captured (the later steps do not refresh it); a step issued after
    MaxIterations is a no-op on the device and must not touch the store either."""
    from gocfd_b200 import lib
    c = Euler(InputParameters2D(CFL=1.0, FluxType="lax", InitType="IVortex", PolynomialOrder=3, FinalTime=50.0,
                                MaxIterations=4, Gamma=1.4, Minf=0.4), structured_tri_mesh(8, 8))
    p = c.problem
    dev, o = lib.Dfr2d(p), ora.OracleSolver(p)
    dev.set_state(c.Q), o.set_state(c.Q)
    dev.capture_edge_values(True)
    dev.step(2), o.step(2)
    _check_gradients(_gradient_fields(dev, p), o)
    edge_rows = dev.gradient_field(200)
    dev.capture_edge_values(False)
    dev.step(1), o.step(1)
    stale_store = ora.OracleSolver(p)                    # oracle with the state of step 3 but the store of step 2
    stale_store.set_state(o.get_state())
    stale_store.EdgeFlux[1][...] = _store_after(p, c.Q, 2)
    _check_gradients(_gradient_fields(dev, p), stale_store)
    assert np.abs(dev.gradient_field(200) - edge_rows).max() > 0
    dev.capture_edge_values(True)
    dev.step(1), o.step(1)                               # step 4 = MaxIterations: finished
    _check_gradients(_gradient_fields(dev, p), o)
    dev.step(3)                                          # no-ops
    _check_gradients(_gradient_fields(dev, p), o)
    dev.close()


def _store_after(p, q0, steps):
    o = ora.OracleSolver(p)
    o.set_state(q0)
    o.step(steps)
    return o.EdgeFlux[1].copy()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4])
def test_device_gradient_fields_with_limiter(n):
    """PerssonC0 path (Q_Face of stage 5 comes from k_diss_prepare; the handle's own element normals are used)."""
    from gocfd_b200 import lib
    c = Euler(InputParameters2D(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n, FinalTime=0.2, MaxIterations=100,
                                Gamma=1.4, Limiter="persson c0", Kappa=5.0), mesh_path("sod-aligned-100pts.su2"))
    p = c.problem
    x, _ = c.DFR.solution_xy()
    w = 0.5 * (1.0 - np.tanh((x - 0.503) / 0.002))
    q0 = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w for v in range(4)])
    dev, o = lib.Dfr2d(p), ora.OracleSolver(p)
    dev.set_state(q0), o.set_state(q0)
    dev.capture_edge_values(True)
    dev.step(3), o.step(3)
    assert o.EpsilonScalar.max() > 0
    _check_gradients(_gradient_fields(dev, p), o, tol=1e-10)
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("diss", [False, True])
def test_device_gradient_fields_partitions_bitwise(diss):
    """Three partitions through dfr2d_multi_step: the neighbour side of a cut edge reads the captured GHOST column, so the
    assembled fields are bitwise those of the single partition."""
    from gocfd_b200 import lib
    kw = dict(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=2, FinalTime=50.0, MaxIterations=100, Gamma=1.4,
              Minf=0.5, Alpha=2.0, LocalTimeStepping=True)
    if diss:
        kw.update(Limiter="persson c0", Kappa=4.0)
    c = Euler(InputParameters2D(**kw), mesh_path("mesh_NACA0012_inv.su2"))
    p = c.problem
    rng = np.random.default_rng(11)
    q = c.Q * (1.0 + 0.03 * rng.standard_normal(c.Q.shape))
    one = lib.Dfr2d(p)
    one.set_state(q)
    one.capture_edge_values(True)
    one.step(2)
    ref = _gradient_fields(one, p)
    devs = [lib.Dfr2d(p, n_parts=3, part=r) for r in range(3)]
    for d in devs:
        d.capture_edge_values(True)
    lib.multi_set_state(devs, q)
    lib.multi_step(devs, 2)
    for pf in GRADIENT_FIELDS:
        out = np.zeros((p.NpFlux, p.K))
        for d in devs:
            d.gradient_field(pf, out)
        np.testing.assert_array_equal(out, ref[pf])
    o = ora.OracleSolver(p)
    o.set_state(q)
    o.step(2)
    _check_gradients(ref, o, tol=1e-10)
    for d in devs + [one]:
        d.close()
