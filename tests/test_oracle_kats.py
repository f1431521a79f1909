"""Pins the host-side operator construction and the oracle against the reference's own
known-answer tests for the hot path.  CPU only.

Each test names the reference test it restates.  Tolerances are the reference's unless a
tighter one is noted.
"""
import math

import numpy as np
import pytest

from conftest import mesh_path
from gocfd_b200.host.dg2d.dfr2d import DFR2D, new_dfr2d
from gocfd_b200.host.euler2d import Euler, PartitionMap
from gocfd_b200.host.input_parameters import InputParameters2D
from oracle.euler2d_oracle import OracleSolver, ivortex_state


def ip_default(**kw):
    # euler_test.go:22-30
    ip = InputParameters2D(CFL=2.5, FluxType="Roe", InitType="Freestream", Minf=2, Gamma=1.4,
                           Limiter="PerssonC0", Kappa=3)
    for k, v in kw.items():
        setattr(ip, k, v)
    return ip


def inviscid_rhs_pieces(c, flux_calc=None, q=None, q_face=None):
    """SetRTFluxInternal -> InterpolateSolutionToEdges -> CalculateEdgeEulerFlux ->
    SetRTFluxOnEdges -> DivInt.Mul (euler_test.go:184-224)."""
    c.problem.Dissipation = False
    o = OracleSolver(c.problem, flux_calc=flux_calc)
    o.set_state(c.Q if q is None else q)
    o.set_rt_flux_internal(o.Q[0])
    if q_face is None:
        o.interpolate_to_edges(o.Q[0])
    else:
        o.Q_Face[...] = q_face
    o.calculate_edge_euler_flux(0.0)
    o.set_rt_flux_on_edges()
    div = np.stack([c.problem.DivInt @ o.F_RT_DOF[n] for n in range(4)])
    return o, div / c.problem.Jdet[None, None, :]


def test_fluid_functions():
    """TestFluidFunctions euler_test.go:32-45."""
    c = Euler(ip_default(Minf=2.0, PolynomialOrder=1), mesh_path("test_tris_6.neu"))
    rho, ru, rv, e = c.Q[:, 0, 0]
    u2 = (ru / rho) ** 2 + (rv / rho) ** 2
    p = 0.4 * (e - 0.5 * rho * u2)
    mach = math.sqrt(u2) / math.sqrt(1.4 * p / rho)
    np.testing.assert_allclose([rho, ru, rv, e, mach, p], [1, 2, 0, 3.78571, 2, 0.71429], atol=1e-5)


@pytest.mark.parametrize("n", range(5))
def test_edge_interpolation_reproduces_constants(n):
    """TestEuler part 1 euler_test.go:55-90."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=n), mesh_path("test_tris_5.neu"))
    kk = c.DFR.K
    q = np.tile(np.arange(1, kk + 1, dtype=np.float64), (c.problem.NpInt, 1))
    qf = c.problem.FluxEdgeInterp @ q
    np.testing.assert_allclose(qf, np.tile(np.arange(1, kk + 1), (3 * c.problem.NpEdge, 1)), atol=1e-6)


@pytest.mark.parametrize("n", range(5))
def test_flux_transformed_layout(n):
    """TestEuler part 2 euler_test.go:93-173: first NpInt rows Fr, next NpInt rows Fs."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=n), mesh_path("test_tris_5.neu"))
    p = c.problem
    kk = p.K
    mark = np.arange(1, kk + 1, dtype=np.float64)[None, :] * np.ones((p.NpInt, 1))
    q = np.stack([mark, 0.1 * mark, 0.05 * mark, 2.0 * mark])
    p.Dissipation = False
    o = OracleSolver(p)
    o.set_state(q)
    o.set_rt_flux_internal(o.Q[0])
    for k in range(kk):
        rho, ru, rv, e = q[:, 0, k]
        u, v = ru / rho, rv / rho
        pr = 0.4 * (e - 0.5 * rho * (u * u + v * v))
        fx = np.array([ru, ru * u + pr, ru * v, u * (e + pr)])
        fy = np.array([rv, rv * u, rv * v + pr, v * (e + pr)])
        ji = p.Jinv[k]
        fr = p.Jdet[k] * (ji[0] * fx + ji[1] * fy)
        fs = p.Jdet[k] * (ji[2] * fx + ji[3] * fy)
        for nn in range(4):
            np.testing.assert_allclose(o.F_RT_DOF[nn][:p.NpInt, k], fr[nn], atol=1e-6)
            np.testing.assert_allclose(o.F_RT_DOF[nn][p.NpInt:2 * p.NpInt, k], fs[nn], atol=1e-6)


@pytest.mark.parametrize("n", range(5))
def test_freestream_divergence_is_zero(n):
    """TestEuler part 3 euler_test.go:176-225 (reference tolerance 1e-6; we hold 1e-10)."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=n), mesh_path("test_tris_6_nowall.neu"))
    _, div = inviscid_rhs_pieces(c)
    assert np.abs(div).max() < 1e-10


def _poly_state(x, y):
    # euler_test.go:869-879 (a=b=c=d=1)
    ax, ay = np.abs(x), np.abs(y)
    rho = ax + ay
    return np.stack([rho, x * rho, y * rho,
                     (x * x + y * y) * (ax / 2.0 + ay / 2.0) + np.power(rho, 1.4) / 0.4])


def _poly_div0(x, y):
    # density component of GetDivergencePoly, euler_test.go:889
    ax, ay = np.abs(x), np.abs(y)
    return (ax + ay) + (ax + ay) + x * (x / ax) + y * (y / ay)


@pytest.mark.parametrize("n", [2, 3, 4])
@pytest.mark.parametrize("mesh", ["test_tris_1tri.neu", "test_tris_two.neu", "test_tris_twoR.neu",
                                  "test_tris_6_nowall.neu"])
def test_polynomial_divergence(n, mesh):
    """CheckFlux0 euler_test.go:903-982: mock flux (rhoU, rhoV) for every equation, analytic
    edge values, divergence vs analytic to 1e-4 relative."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=n), mesh_path(mesh))
    p = c.problem
    x, y = c.DFR.flux_xy()
    qflux = _poly_state(x, y)
    q = qflux[:, :p.NpInt]
    qface = qflux[:, 2 * p.NpInt:]
    mock = lambda rho, ru, rv, e: ((ru, ru, ru, ru), (rv, rv, rv, rv))     # noqa: E731
    _, div = inviscid_rhs_pieces(c, flux_calc=mock, q=q, q_face=qface)
    want = _poly_div0(x[:p.NpInt], y[:p.NpInt])
    for nn in range(4):
        np.testing.assert_allclose(div[nn] / q[0], want / q[0], atol=1e-4)


def test_vortex_initial_condition_and_divergence():
    """TestEuler part 4 euler_test.go:258-323: IC equals analytic state (1e-6); density RHS vs
    analytic divergence within 1e-3 relative at N=1."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=1, InitType="ivortex"), mesh_path("test_tris_6.neu"))
    t = c.DFR.Tris
    t.bcType[t.bcType == 9] = 0          # BC_IVortex -> BC_None, euler_test.go:267-271
    c.problem.edge_bc = t.bcType.copy()
    p = c.problem
    x, y = c.DFR.solution_xy()
    st = ivortex_state((5.0, 5.0, 0.0, 1.4, 1.0), 0.0, x, y)
    for nn in range(4):
        np.testing.assert_allclose(c.Q[nn], st[nn], atol=1e-6)
    _, div = inviscid_rhs_pieces(c)
    # analytic density-equation divergence by differentiating the analytic mass flux numerically
    h = 1e-6
    def mass_flux(xx, yy):
        s = ivortex_state((5.0, 5.0, 0.0, 1.4, 1.0), 0.0, xx, yy)
        return s[1], s[2]
    dfx = (mass_flux(x + h, y)[0] - mass_flux(x - h, y)[0]) / (2 * h)
    dgy = (mass_flux(x, y + h)[1] - mass_flux(x, y - h)[1]) / (2 * h)
    np.testing.assert_allclose(div[0] / st[0], (dfx + dgy) / st[0], atol=1e-3)


def test_vortex_centre_state():
    """TestIVortex isentropic_vortex/analytic_vortex_test.go:11-35 (centre state)."""
    st = ivortex_state((5.0, 5.0, 0.0, 1.4, 1.0), 0.0, np.array([5.0]), np.array([0.0]))
    np.testing.assert_allclose([s[0] for s in st], [0.361673, 0.361673, 0.0, 0.782817], atol=1e-5)


def test_vertex_to_element_and_max_merge():
    """TestDissipation euler_test.go:418-518: element areas, vertex grouping, max-merge vector."""
    dfr = new_dfr2d(1, mesh_path("test_tris_9.neu"))
    assert dfr.K == 10
    np.testing.assert_allclose(2.0 * dfr.Jdet, 0.25, atol=1e-6)
    want = [(0, 0), (0, 1), (1, 3), (1, 1), (1, 2), (2, 4), (2, 3), (3, 5), (3, 0), (4, 2), (4, 5),
            (4, 6), (4, 1), (4, 0), (4, 7), (5, 4), (5, 3), (5, 9), (5, 8), (5, 2), (5, 7), (6, 4),
            (6, 9), (7, 5), (7, 6), (8, 8), (8, 6), (8, 7), (9, 8), (9, 9)]
    have = sorted((int(v), k) for k in range(dfr.K) for v in dfr.EToV[k])
    assert have == sorted(want)
    ip = ip_default(PolynomialOrder=1, FluxType="average")
    c = Euler(ip, mesh_path("test_tris_9.neu"))
    assert c.problem.Dissipation
    o = OracleSolver(c.problem)
    o.EpsilonScalar = np.arange(10, dtype=np.float64)
    o.SigmaScalar = np.arange(10, dtype=np.float64)
    o.merge_to_vertices()
    assert list(o.EpsVertex) == [1, 3, 4, 5, 7, 9, 9, 6, 8, 9]


def _poly_q_and_grads(x, y):
    q = [x + y, x * x + y * y, x ** 3 + y ** 3, x ** 4 + y ** 4]
    gx = [np.ones_like(x), 2 * x, 3 * x * x, 4 * x ** 3]
    gy = [np.ones_like(y), 2 * y, 3 * y * y, 4 * y ** 3]
    return q, gx, gy


def test_rt_gradient_of_polynomials():
    """TestDissipation2 euler_test.go:565-687: Div.(DXMetric*U) is the exact x-derivative of
    x^m + y^m, m=1..4, at N=4 on test_tris_5 (1e-6), both hand-rolled and through the oracle's
    GetSolutionGradientUsingRTElement restatement."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=4), mesh_path("test_tris_5.neu"))
    p = c.problem
    dfr = c.DFR
    x, y = dfr.flux_xy()
    q, gx, gy = _poly_q_and_grads(x, y)
    dxm, dym = dfr.metrics()
    ni = p.NpInt
    for n in range(4):
        un = np.concatenate([q[n][:ni], q[n][:ni], p.FluxEdgeInterp @ q[n][:ni]])
        np.testing.assert_allclose(p.Div @ (dxm * un), gx[n], atol=1e-6)
        np.testing.assert_allclose(p.Div @ (dym * un), gy[n], atol=1e-6)
    o = OracleSolver(p)
    qs = np.stack([q[n][:ni] for n in range(4)])
    o.set_state(qs)
    o.interpolate_to_edges(o.Q[0])
    for n in range(4):
        o.EdgeFlux[1][n] = o.Q_Face[n][o.rowsL, o.kLc]
    o.Epsilon[...] = 1.0
    o.calculate_epsilon_gradient(o.Q[0])
    for n in range(4):
        np.testing.assert_allclose(o.DissX[n], gx[n], atol=1e-6)
        np.testing.assert_allclose(o.DissY[n], gy[n], atol=1e-6)


def test_freestream_gradient_is_zero():
    """TestEuler_GetSolutionGradientUsingRTElement euler_test.go:690-748 (N=4, test_tris_9)."""
    c = Euler(ip_default(FluxType="average", PolynomialOrder=4, Minf=0.8, Alpha=2.0), mesh_path("test_tris_9.neu"))
    o = OracleSolver(c.problem)
    o.set_state(c.Q)
    o.interpolate_to_edges(o.Q[0])
    o.calculate_edge_euler_flux(0.0)
    o.Epsilon[...] = 1.0
    o.calculate_epsilon_gradient(o.Q[0])
    assert np.abs(o.DissX).max() < 1e-6 and np.abs(o.DissY).max() < 1e-6


def test_edges_of_test_tris_9():
    """TestEdges edges_test.go:71-96: 19 edges and the two vertex orderings."""
    dfr = new_dfr2d(1, mesh_path("test_tris_9.neu"))
    keys = dfr.Tris.key
    assert len(keys) == 19
    hi = sorted(int(k >> np.uint64(32)) for k in keys)
    lo = sorted(int(k & np.uint64(0xFFFFFFFF)) for k in keys)
    assert hi == [1, 2, 3, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 8, 9, 9, 9]
    assert lo == [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 4, 5, 5, 5, 6, 7, 8]


def test_dfr2d_edge_table_and_normals():
    """TestDFR2D DG2D/dfr_startup_test.go:17-204: shared edge (0,2) is tri0 edge 2 / tri1 edge 0;
    EtoE; |edge|-scaled normals of element 0 are {0,-1},{10,4}->... checked through IInII."""
    dfr = new_dfr2d(1, mesh_path("test_tris_5.neu"))
    t = dfr.Tris
    shared = np.flatnonzero(t.nConn == 2)
    assert len(shared) == 1
    e = shared[0]
    assert (t.kL[e], t.edgeNumL[e], t.kR[e], t.edgeNumR[e]) == (0, 2, 1, 0)
    assert t.EtoE.tolist() == [[-1, -1, 1], [0, -1, -1]]
    # the neighbour's normal is exactly the negative of the owner's
    assert dfr.FaceNorm[0, 2, 0] == -dfr.FaceNorm[0, 0, 1]
    assert dfr.FaceNorm[1, 2, 0] == -dfr.FaceNorm[1, 0, 1]
    # ||n|| scaling: |edge|/2 on edges 0 and 2, |edge|/(2 sqrt 2) on the hypotenuse
    vx, vy = dfr.VX, dfr.VY
    ev = dfr.EToV[0]
    for le in range(3):
        a, b = ev[le], ev[(le + 1) % 3]
        length = math.hypot(vx[b] - vx[a], vy[b] - vy[a])
        want = length / 2.0 if le != 1 else length / (2.0 * math.sqrt(2.0))
        assert abs(dfr.IInII[le, 0] - want) < 1e-12
    # outward normals: n . (edge midpoint - centroid) > 0
    for k in range(dfr.K):
        ev = dfr.EToV[k]
        cx, cy = vx[ev].mean(), vy[ev].mean()
        for le in range(3):
            a, b = ev[le], ev[(le + 1) % 3]
            mx, my = 0.5 * (vx[a] + vx[b]), 0.5 * (vy[a] + vy[b])
            assert dfr.FaceNorm[0, le, k] * (mx - cx) + dfr.FaceNorm[1, le, k] * (my - cy) > 0
    # BC: Inflow on element 1 face 1 and element 2 face 2 (test_tris_5.neu)
    assert sorted(t.bcType[t.nConn == 1].tolist()) == [0, 0, 1, 1]


@pytest.mark.parametrize("n", [2, 3, 4])
def test_shock_finder_pattern(n):
    """TestShockFinder DG2D/dfr_shock_capturing_test.go:11-50: Mach-2 normal shock at x=0 on
    test_10tris_centered; ElementHasShock = {F,F,T,T,F,F,F,T,T,F}, Kappa=3, threshold 0.0075,
    ShockIndicator variant (S0 = 4/N^4)."""
    dfr = new_dfr2d(n, mesh_path("test_10tris_centered.neu"))
    x, _ = dfr.solution_xy()
    rho = np.where(x < 0, 1.0, 8.0 / 3.0)
    sf = dfr.shock_finder()
    mass = np.diag(sf.MassMatrix)[:, None]
    clipped = sf.Clipper @ rho
    m = (mass * (rho - clipped) ** 2).sum(axis=0) / (mass * rho * rho).sum(axis=0)
    with np.errstate(divide="ignore"):
        se = np.log10(m)
    kappa, s0 = 3.0, 4.0 / float(n) ** 4
    sigma = np.where(se < s0 - kappa, 0.0,
                     np.where(se <= s0 + kappa, 0.5 * (1 + np.sin(math.pi * (0.5 / kappa) * (se - s0))), 1.0))
    f, t = False, True
    assert (sigma > 0.0075).tolist() == [f, f, t, t, f, f, f, t, t, f]
    # D and P used by the RK-path sensor are consistent with the clipper
    np.testing.assert_allclose(sf.D @ rho, rho - clipped, atol=1e-13)
    np.testing.assert_allclose(sf.P, sf.MassMatrix @ sf.D, atol=0)


def test_mass_matrix_definition():
    """lagrange_element.go:78-82: MassMatrix = V^T diag(W) V (symmetric; its leading entry is the
    exact integral of the constant mode, 1/2 with weights normalised to 1)."""
    for n in range(5):
        el = DFR2D(n).SolutionElement
        np.testing.assert_allclose(el.MassMatrix, el.MassMatrix.T, atol=1e-14)
        assert abs(el.MassMatrix[0, 0] - 0.5) < 1e-12


def test_partition_map():
    """TestEuler_Indexing utils/parallel_utils_test.go:11-63."""
    def histo(k, npar):
        pm = PartitionMap(npar, k)
        h = {}
        for b in range(npar):
            d = pm.get_bucket_dimension(b)
            h[d] = h.get(d, 0) + 1
        return h
    assert histo(2, 32) == {0: 30, 1: 2}
    assert histo(32, 32) == {1: 32}
    assert histo(256, 32) == {8: 32}
    assert histo(287, 32) == {8: 1, 9: 31}
    for n in range(64, 2000, 7):
        h = histo(n, 32)
        assert sum(k * c for k, c in h.items()) == n
        if len(h) == 2:
            a, b = h.keys()
            assert abs(a - b) == 1
    for max_index in range(10, 300, 13):
        pm = PartitionMap(5, max_index)
        for k in range(max_index):
            bn, lo, hi = pm.get_bucket(k)
            assert lo <= k < hi and (lo, hi) == pm.get_bucket_range(bn)


def test_su2_reader_naca():
    """TestReadSU2 readfiles/readSU2_test.go:11 (shape checks on the shipped NACA mesh)."""
    from gocfd_b200.host.readfiles import read_su2
    m = read_su2(mesh_path("mesh_NACA0012_inv.su2"))
    assert m.K == 10216 and len(m.VX) == 5233
    assert {k: len(v) for k, v in m.BCEdges.items()} == {"wall": 200, "far": 50}


# ---- DG2D/dfr_startup_test.go:224-247, :293-368 (TestDivergence) and :370-489 (TestGradient) ---------------------

def _equi_tri_mesh(angle):
    """CreateEquiTriMesh: one equilateral triangle of side 5, rotated by `angle` degrees."""
    from gocfd_b200.host.readfiles import Mesh2D
    scale = 5.0
    y_height = scale * math.sin(math.pi * 60.0 / 180.0)
    vx = np.array([-scale * 0.5, scale * 0.5, 0.0])
    vy = np.array([-y_height / 3.0, -y_height / 3.0, 2.0 * y_height / 3.0])
    ca, sa = math.cos(2.0 * math.pi * angle / 360.0), math.sin(2.0 * math.pi * angle / 360.0)
    return Mesh2D(vx * ca + vy * sa, -vx * sa + vy * ca, [[0, 1, 2]], {})


def _project_flux_onto_rt_space(dfr, fx, fy):
    """ProjectFluxOntoRTSpace (DG2D/dfr_startup.go:316-358): [NpFlux, K]."""
    rt = dfr.FluxElement
    ni, ne = rt.NpInt, rt.NpEdge
    ji, jd = dfr.Jinv, dfr.Jdet
    ft0 = jd * (ji[:, 0] * fx + ji[:, 1] * fy)
    ft1 = jd * (ji[:, 2] * fx + ji[:, 3] * fy)
    fp = np.empty_like(fx)
    fp[:ni] = ft0[:ni]
    fp[ni:2 * ni] = ft1[ni:2 * ni]
    e0 = 2 * ni
    fp[e0:e0 + ne] = -ft1[e0:e0 + ne]
    fp[e0 + ne:e0 + 2 * ne] = (1.0 / math.sqrt(2.0)) * (ft0[e0 + ne:e0 + 2 * ne] + ft1[e0 + ne:e0 + 2 * ne])
    fp[e0 + 2 * ne:] = -ft0[e0 + 2 * ne:]
    return fp


def _divergence_check(dfr):
    rt = dfr.FluxElement
    ni, ne = rt.NpInt, rt.NpEdge
    x, y = dfr.flux_xy()
    for order in range(1, dfr.N + 1):
        fx, fy = x ** order, y ** order
        want = order * (x ** (order - 1) + y ** (order - 1))
        tol = np.abs(fx).max() * 1.0e-9
        fp = _project_flux_onto_rt_space(dfr, fx, fy)
        assert np.abs(rt.Div @ fp / dfr.Jdet - want).max() <= tol
        # SetNormalFluxOnEdges: physical unit normal scaled by ||n|| (IInII) on the edge rows, same divergence
        for e in range(3):
            rows = slice(2 * ni + e * ne, 2 * ni + (e + 1) * ne)
            fp[rows] = dfr.IInII[e] * (dfr.FaceNorm[0, e] * fx[rows] + dfr.FaceNorm[1, e] * fy[rows])
        assert np.abs(rt.Div @ fp / dfr.Jdet - want).max() <= tol


@pytest.mark.parametrize("angle", [0, 15, 25, 45, 90, 180, 210, 270, 310])
def test_rt_divergence_rotated_equilateral(angle):
    """TestDivergence: RT divergence of (x^m, y^m), m = 1..4, is exact at N=4 on a rotated equilateral triangle,
    through the Piola projection and through the IInII-scaled physical normals (tolerance 1e-9 max|Fx|)."""
    _divergence_check(DFR2D(4, _equi_tri_mesh(angle)))


def test_rt_divergence_test_tris_5():
    _divergence_check(new_dfr2d(4, mesh_path("test_tris_5.neu")))


def test_gradient_on_flux_points():
    """TestGradient (N=3, test_tris_6.neu): d/dx, d/dy of x+y, x^2+y^2, x^3+y^3 at the RT points, (a) through
    FluxDr/FluxDs + Jinv and (b) through the RT element, Div . (DXMetric (.) U) with U interpolated to the edges by
    FluxEdgeInterp -- the operator chain of GetSolutionGradientUsingRTElement.  Tolerance 1e-6."""
    dfr = new_dfr2d(3, mesh_path("test_tris_6.neu"))
    assert dfr.K == 2
    rt = dfr.FluxElement
    ni = rt.NpInt
    x, y = dfr.flux_xy()
    dxm, dym = dfr.metrics()
    for m in (1, 2, 3):
        u_int = (x ** m + y ** m)[:ni]
        dx_want, dy_want = m * x ** (m - 1), m * y ** (m - 1)
        qr, qs = dfr.FluxDr @ u_int, dfr.FluxDs @ u_int
        ji = dfr.Jinv
        np.testing.assert_allclose(qr * ji[:, 0] + qs * ji[:, 2], dx_want, atol=1e-6)
        np.testing.assert_allclose(qr * ji[:, 1] + qs * ji[:, 3], dy_want, atol=1e-6)
        un = np.concatenate([u_int, u_int, dfr.FluxEdgeInterp @ u_int])
        np.testing.assert_allclose(rt.Div @ (dxm * un), dx_want, atol=1e-6)
        np.testing.assert_allclose(rt.Div @ (dym * un), dy_want, atol=1e-6)


# ---- DG2D/raviart_thomas_element_test.go:17-163 (TestRTElement) -----------------------------------------------

@pytest.mark.parametrize("p_rt", [1, 2, 3, 4, 5])
def test_rt_element_polynomial_divergence(p_rt):
    """DivergencePolynomialField_Test: for RT order P = N+1 the divergence operator `Div` applied to the DOF projection
    of the three reference polynomial fields of order 0..P equals the analytic divergence at every RT point
    (tolerance 1e-9 max|f1|, the reference's)."""
    from gocfd_b200.host.dg2d.elements import RTElement
    rt = RTElement(p_rt)
    r, s = rt.R, rt.S
    dof = np.asarray(rt.DOFVectors)
    fields = [
        (lambda p: ((r + s + 10) ** p, (10 * (r + s)) ** p),
         lambda p: p * ((r + s + 10) ** (p - 1) + 10 * (10 * (r + s)) ** (p - 1)) if p > 0 else 0 * r),
        (lambda p: (r ** p, s ** p),
         lambda p: p * (r ** (p - 1) + s ** (p - 1)) if p > 0 else 0 * r),
        (lambda p: ((s + 10) ** p, (10 * r) ** p),
         lambda p: 0 * r),
    ]
    for f, div in fields:
        for p in range(0, p_rt + 1):
            f1, f2 = f(p)
            proj = dof[:, 0] * f1 + dof[:, 1] * f2          # ProjectFunctionOntoDOF
            max_f = np.abs(f1).max()
            tol = 1e-9 * max_f if max_f >= 1e-9 else 1e-9
            assert np.abs(rt.Div @ proj - div(p)).max() <= tol


def test_su2_reader_reference_fixture():
    """TestReadSU2 (readfiles/readSU2_test.go:11-72) on the reference's own inline 22-element gmsh export
    (tests/golden/meshes/su2_reader_kat.su2 is that literal): element/vertex counts, last element, last vertex to the
    last bit, markers in file order with the two "periodic-rightleft" blocks concatenated."""
    from gocfd_b200.host.readfiles import read_su2
    m = read_su2(mesh_path("su2_reader_kat.su2"))
    assert m.K == 22 and int(m.EToV[m.K - 1, 2]) == 17
    assert len(m.VX) == 18 and len(m.VY) == 18
    assert m.VX[-1] == -7.100939331382065 and m.VY[-1] == 2.889910324036197
    assert list(m.BCEdges) == ["periodic-rightleft", "top", "bottom"]
    assert [len(v) for v in m.BCEdges.values()] == [4, 4, 4]          # 2 + 2 periodic, 4 top, 4 bottom
    bottom, per = m.BCEdges["bottom"], m.BCEdges["periodic-rightleft"]
    assert bottom[2].tolist() == [5, 6] and bottom[3].tolist() == [6, 1]
    assert per.tolist() == [[3, 11], [11, 0], [1, 7], [7, 2]]


def test_input_parameters_parse():
    """TestInputParameters_Parse (model_problems/Euler2D/euler_test.go:750-776), same document and checks."""
    doc = """
Title: Test Case
CFL: 1.
InitType: Freestream # Can be IVortex or Freestream
FluxType: Roe
PolynomialOrder: 2
FinalTime: 4.
BCs: 
  Inflow:
      37:
         NPR: 4.0
  Outflow:
      22:
         P: 1.5
"""
    ip = InputParameters2D().parse(doc)
    assert ip.BCs["Inflow"][37]["NPR"] == 4.0
    assert ip.BCs["Outflow"][22]["P"] == 1.5
    assert ip.FinalTime == 4.0 and ip.CFL == 1.0 and ip.PolynomialOrder == 2
    assert ip.InitType == "Freestream" and ip.FluxType == "Roe" and ip.Title == "Test Case"


def test_element_mean_of_mach5_shock_field():
    """TestElementMean (model_problems/Euler2D/dfr_shock_capturing_Euler_test.go:15-56; runs under `go test -v`):
    Williams-Shunn-Jameson weighted element means (UpdateElementMean, euler.go:1004-1020) of the NORMALSHOCKTESTM5 field
    (DG2D/test_functions2.go:133-236: u1 = 5, p1 = rho1 = 1, Rankine-Hugoniot state for x >= 0) at N=2 on
    test_10tris_centered.neu.  Pins the solution-point coordinates and the cubature weights that build the mass matrix."""
    dfr = new_dfr2d(2, mesh_path("test_10tris_centered.neu"))
    g = 1.4
    m1n = 5.0 / math.sqrt(g)
    rho_ratio = ((g + 1) * m1n * m1n) / ((g - 1) * m1n * m1n + 2)
    p_ratio = 1 + (2 * g / (g + 1)) * (m1n * m1n - 1)
    u2 = 5.0 / rho_ratio
    U1 = [1.0, 5.0, 0.0, 1.0 / (g - 1) + 0.5 * 25.0]
    U2 = [rho_ratio, rho_ratio * u2, 0.0, p_ratio / (g - 1) + 0.5 * rho_ratio * u2 * u2]
    x, _ = dfr.solution_xy()
    w = dfr.SolutionElement.W
    assert abs(w.sum() - 1.0) < 1e-12
    means = [(w[:, None] * np.where(x < 0, U1[n], U2[n])).sum(axis=0) for n in range(4)]
    np.testing.assert_allclose(means[0], [1, 1, 1.40545, 4.28205, 4.6875, 1, 1, 1.40545, 4.28205, 4.6875], atol=1e-4)
    np.testing.assert_allclose(means[1], [5] * 10, atol=1e-4)
    np.testing.assert_allclose(means[2], [0] * 10, atol=1e-4)
    np.testing.assert_allclose(means[3], [15, 15, 19.32477, 50.00856, 54.33333, 15, 15, 19.32477, 50.00856, 54.33333],
                               atol=1e-4)
