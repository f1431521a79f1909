"""Host-side post-processors of Euler.OutputFinal: analytic Sod solution against the
reference's own known-answer test, centre-line interpolation, wall plot data.  CPU only."""
import numpy as np
import pytest

from conftest import mesh_path
from host_standin import output_final as of
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from host_standin.sod_shock_tube import SODExact, SODShockTube, shocktube_files


def test_sod_exact_reference_kat():
    """TestSOD (model_problems/Euler1D/sod_shock_tube/analytic_sod_test.go:17-31), same vectors and tolerances."""
    sod = SODExact(0.1)
    x, rho, _, _, _ = sod.get()
    x_check = [0, 0.3815784043380077, 0.3817784043380077, 0.39280783577858336, 0.40393726721915907,
               0.4150666986597348, 0.42619613010031043, 0.4373255615408861, 0.4484549929814618, 0.4595844244220375,
               0.47071385586261316, 0.4818432873031888, 0.49277271874376455, 0.49287271874376454, 0.4930727187437645,
               0.5926452620047974, 0.5928452620047974, 0.675115573202932, 0.675315573202932, 1]
    rho_check = [1, 1, 0.9992959031724784, 0.9240353444481086, 0.852758969991083, 0.7859504402212434,
                 0.7233963393812908, 0.6648901587403833, 0.6102321829702019, 0.5592293765210307, 0.5116952699978237,
                 0.467449846536279, 0.4270320564069276, 0.42667562327066666, 0.4263194281781805, 0.4263194281781805,
                 0.26557371170513905, 0.26557371170513905, 0.125, 0.125]
    assert len(x) == 20
    assert np.abs(x - x_check).max() < 0.001 and np.abs(rho - rho_check).max() < 0.001
    # the restatement follows the same arithmetic, so it actually agrees to round-off of the secant iteration
    assert np.abs(x - x_check).max() < 1e-12 and np.abs(rho - rho_check).max() < 1e-12
    assert abs(sod.x4 - 0.6752) < 0.0001
    assert abs(SODExact(0.2).x4 - 0.8504) < 0.0001


def test_sod_exact_jump_conditions():
    """Rankine-Hugoniot across the shock and constant p, u across the contact."""
    s = SODExact(0.2)
    rl, pl, ul, _, _ = s.getx(s.x4 - 1e-6)
    rr, pr, ur, _, _ = s.getx(s.x4 + 1e-6)
    vs = (s.x4 - s.x0) / s.t
    assert abs(rl * (ul - vs) - rr * (ur - vs)) < 1e-6                       # mass (root tolerance 1e-7 on f)
    assert abs((rl * (ul - vs) ** 2 + pl) - (rr * (ur - vs) ** 2 + pr)) < 1e-6   # momentum
    a, b = s.getx(s.x3 - 1e-6), s.getx(s.x3 + 1e-6)
    assert a[1] == b[1] and a[2] == b[2] and a[0] != b[0]


def _sod_case(n=2):
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n, FinalTime=0.2,
                           MaxIterations=10, Gamma=1.4, Limiter="persson c0", Kappa=5.0)
    return Euler(ip, mesh_path("sod-aligned-100pts.su2"))


def test_shock_tube_sampling_and_interpolation():
    c = _sod_case(2)
    st = SODShockTube(4 * c.DFR.K // 5, c.DFR)
    assert st.Npts == 4 * c.DFR.K // 5
    assert st.XLocations[0] == pytest.approx(1e-5) and st.XLocations[-1] == pytest.approx(1 - 1e-5)
    # every target lies in its element: barycentric weights of getUVCoords
    assert (st.RS >= 0).all() and (st.RS.sum(axis=1) <= 1 + 1e-15).all()
    dfr = c.DFR
    ymid = 0.5 * (dfr.VY.max() - dfr.VY.min()) + dfr.VY.min()
    for i in (0, st.Npts // 3, st.Npts - 1):
        k = st.ElementNumber[i]
        a, b, cc = dfr.EToV[k]
        r, s = st.RS[i]
        x = dfr.VX[a] + r * (dfr.VX[cc] - dfr.VX[a]) + s * (dfr.VX[b] - dfr.VX[a])
        y = dfr.VY[a] + r * (dfr.VY[cc] - dfr.VY[a]) + s * (dfr.VY[b] - dfr.VY[a])
        assert x == pytest.approx(st.XLocations[i], abs=1e-12) and y == pytest.approx(ymid, abs=1e-12)
    # interpolation rows reproduce constants exactly (rows of a Lagrange interpolation sum to one) ...
    assert np.abs(st.InterpolationMatrix.sum(axis=1) - 1.0).max() < 1e-11
    q = np.empty_like(c.Q)
    for n in range(4):
        q[n] = 1.0 + n
    st.interpolate_fields(q)
    np.testing.assert_allclose(st.Rho, 1.0, rtol=1e-11)
    np.testing.assert_allclose(st.RhoU, 2.0, rtol=1e-11)
    np.testing.assert_allclose(st.E, 4.0, rtol=1e-11)
    # ... and on the initial condition the samples are the left / right states away from the split
    st.interpolate_fields(c.Q)
    assert st.Rho[0] == pytest.approx(1.0) and st.Rho[-1] == pytest.approx(0.125)
    assert st.E[0] == pytest.approx(2.5) and st.E[-1] == pytest.approx(0.25)
    # and it is the statement of InterpolateFields: row . column
    i = st.Npts // 4
    assert st.Rho[i] == pytest.approx(float(st.InterpolationMatrix[i] @ c.Q[0][:, st.ElementNumber[i]]), rel=1e-14)
    num, ana = shocktube_files(st, "sod.su2", 0.2)
    lines = num.splitlines()
    assert lines[0] == "Meshfile: sod.su2" and lines[1] == "X\tRho\tRhoU\tE" and len(lines) == 2 + st.Npts
    assert len(lines[2].split("\t")) == 4 and lines[2].startswith("0.00001000\t1.00000000\t")
    assert len(ana.splitlines()) == 21


def test_best_match_flow_function():
    assert of.best_match_flow_function("Pressure Coefficient") == (7, True)
    assert of.best_match_flow_function("pressure_coefficient") == (7, True)
    assert of.best_match_flow_function("Mach") == (4, True)
    assert of.best_match_flow_function("density")[1]
    assert of.best_match_flow_function("Entropy") == (13, True)
    assert of.best_match_flow_function("zzz")[1] is False


def test_wall_plot_data_freestream_naca(tmp_path):
    """plotfile.dat rows: wall edge points lie on the airfoil surface; Cp of the undisturbed freestream is 0 and
    Mach is Minf at every wall point; row count = wall edges x NpEdge."""
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=2, FinalTime=1.0,
                           MaxIterations=10, Gamma=1.4, Minf=0.8, Alpha=2.0, PlotFields=["Pressure Coefficient", "Mach"])
    c = Euler(ip, mesh_path("mesh_NACA0012_inv.su2"))
    text, (n_wall, n_pts) = of.wall_plot_data(c, c.Q)
    assert (n_wall, n_pts) == (200, 4)
    rows = [[float(v) for v in ln.rstrip(",").split(",")] for ln in text.splitlines()]
    assert len(rows) == 800 and all(len(r) == 4 for r in rows)
    a = np.array(rows)
    assert a[:, 0].min() >= -1e-6 and a[:, 0].max() <= 1.01          # chord [0, 1]
    assert np.abs(a[:, 1]).max() < 0.07                               # NACA 0012 half thickness 0.06
    assert np.abs(a[:, 2]).max() < 1e-9                               # Cp of the freestream
    np.testing.assert_allclose(a[:, 3], 0.8, rtol=1e-5)               # %.5e columns
    # a perturbed state: the wall value equals GraphInterp . f(Q) at the edge point of the graph mesh
    rng = np.random.default_rng(3)
    q = c.Q * (1.0 + 0.01 * rng.standard_normal(c.Q.shape))
    t = c.DFR.Tris
    e = int(np.nonzero(t.bcType == of.BC_Wall)[0][7])
    k, en = int(t.kL[e]), int(t.edgeNumL[e])
    f = of.plot_field_elements(c, q, 7, np.array([k]))[0]
    cp_nodes = of.get_flow_function(c.FSFar, [q[n][:, [k]] for n in range(4)], 7)[:, 0]
    ned = c.DFR.FluxElement.NpEdge
    want = c.DFR.FluxEdgeInterp[en * ned:(en + 1) * ned] @ cp_nodes      # same points, same basis
    np.testing.assert_allclose(f[en * (ned + 1) + 1:en * (ned + 1) + 1 + ned], want, rtol=1e-10, atol=1e-12)
    # OutputFinal writes the file and prints the reference's summary line
    msgs = []
    paths = of.output_final(c, c.Q, outdir=str(tmp_path), out=msgs.append)
    assert [p.split("/")[-1] for p in paths] == ["plotfile.dat"]
    assert msgs == ["Output plot data for wall, dimensions: 200 Wall edges by 4 points each"]


def test_output_final_shocktube_files(tmp_path):
    c = _sod_case(1)
    paths = of.output_final(c, c.Q, mesh_file="sod-aligned-100pts.su2", outdir=str(tmp_path))
    assert [p.split("/")[-1] for p in paths] == ["shocktube.dat", "shocktube_analytic.dat"]
    assert open(paths[0]).readline() == "Meshfile: sod-aligned-100pts.su2\n"


def test_solve_writes_output_final(tmp_path):
    """Euler.Solve(..., output_dir=...) ends like the reference's Solve: PrintFinal, then OutputFinal (euler.go:217-219).
    Driven by the oracle here (CPU); the device library has the same solver surface."""
    from oracle.euler2d_oracle import OracleSolver
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=1, FinalTime=0.2,
                           MaxIterations=3, Gamma=1.4, Limiter="persson c0", Kappa=5.0)
    c = Euler(ip, mesh_path("sod-aligned-100pts.su2"))
    lines = []
    steps, _ = c.Solve(OracleSolver(c.problem), out=lines.append, output_dir=str(tmp_path), mesh_file="sod-aligned-100pts.su2")
    assert steps == 3
    rows = open(tmp_path / "shocktube.dat").read().splitlines()
    assert rows[0] == "Meshfile: sod-aligned-100pts.su2" and len(rows) == 2 + 4 * c.DFR.K // 5
    vals = np.array([[float(v) for v in r.split("\t")] for r in rows[2:]])
    assert vals[0, 1] == pytest.approx(1.0, abs=1e-6) and vals[-1, 1] == pytest.approx(0.125, abs=1e-6)
    assert (tmp_path / "shocktube_analytic.dat").exists()


def test_solve_progress_lines_follow_the_reference_format():
    """PrintInitialization / PrintUpdate / PrintFinal (euler.go:802-862): header, "%8d%8.5f%8.5f" + six "%11.4e" columns
    (Res0..3 as max(0, signed max), L1 = max, L2 = sqrt(sum sq)/4), and the rate line."""
    import re
    from oracle.euler2d_oracle import OracleSolver
    from gocfd_b200.host.meshgen import structured_tri_mesh
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=1, FinalTime=100.0,
                           MaxIterations=3, Gamma=1.4, Minf=0.1)
    c = Euler(ip, structured_tri_mesh(6, 6, tag="wall"))
    lines = []
    c.Solve(OracleSolver(c.problem), out=lines.append)
    assert lines[0].splitlines()[0] == "Solving until finaltime = 100.00000"
    assert lines[0].splitlines()[1] == "    iter    time      dt       Res0       Res1       Res2       Res3         L1         L2"
    upd = [ln for ln in lines[1:] if re.match(r"^\s+\d+ ", ln)]
    assert len(upd) >= 2                                  # step 1 and the final step
    first = upd[0]
    assert re.fullmatch(r"\s{7}1\s+\d\.\d{5}\s*\d\.\d{5}(\s+\d\.\d{4}e[+-]\d\d){6}", first), first
    cols = [float(x) for x in first.split()[3:]]
    assert cols[4] == pytest.approx(max(cols[:4]), rel=1e-3)
    assert cols[5] == pytest.approx(np.sqrt(sum(v * v for v in cols[:4])) / 4.0, rel=1e-3)
    assert re.search(r"Rate of execution =\s+\d+\.\d{5} us/\(element\*iteration\) over 3 iterations", lines[-1])
    # local time stepping prints the iteration only (euler.go:816-817)
    ip.LocalTimeStepping = True
    c = Euler(ip, structured_tri_mesh(6, 6, tag="wall"))
    lines = []
    c.Solve(OracleSolver(c.problem), out=lines.append)
    assert lines[0].startswith("Solving until Max Iterations = 3\n    iter                ")
    assert lines[1].startswith("         1              ")


def test_c1_run_script_configuration_end_to_end(tmp_path):
    """Config C1 exactly as the reference's run script builds it (test_cases/Euler2D/SU2-naca12/runme.sh:7-13):
    input-base.yaml with the overrides appended, so PolynomialOrder / CFL / MaxIterations appear twice and the last
    one wins; then a (shortened) run and OutputFinal's wall plot file with "mach" and "pressure coefficient"."""
    from oracle.c_oracle import COracleSolver
    base = """Title: "Test Case"
FluxType: Roe
InitType: Freestream # Can be "Freestream" or "IVortex"
FinalTime: 20
Gamma: 1.4
LocalTimeStepping: true
Limiter: PerssonC0
Kappa: 4.5
PlotFields: ["mach", "pressure coefficient"]
PolynomialOrder: 4
CFL: 5
Minf: 0.8
Alpha: 1.25
MaxIterations: 50000
"""
    doc = base + "PolynomialOrder: 0\nCFL: 2.0\nMaxIterations: 2000\n"
    ip = InputParameters2D(Gamma=1.4, Minf=0.1).parse(doc)
    assert (ip.PolynomialOrder, ip.CFL, ip.MaxIterations) == (0, 2.0, 2000)
    assert ip.LocalTimeStepping is True and ip.Limiter == "PerssonC0" and ip.Kappa == 4.5 and ip.Alpha == 1.25
    assert ip.PlotFields == ["mach", "pressure coefficient"]
    assert [of.best_match_flow_function(f) for f in ip.PlotFields] == [(4, True), (7, True)]
    ip.MaxIterations = 20                      # shortened run
    c = Euler(ip, mesh_path("mesh_NACA0012_inv.su2"))
    assert not c.Dissipation                   # N == 0 disables the limiter (euler.go:110)
    lines = []
    steps, _ = c.Solve(COracleSolver(c.problem), out=lines.append, output_dir=str(tmp_path))
    assert steps == 20
    assert lines[-1] == "Output plot data for wall, dimensions: 200 Wall edges by 1 points each"
    rows = [[float(v) for v in ln.rstrip(",").split(",")] for ln in open(tmp_path / "plotfile.dat").read().splitlines()]
    a = np.array(rows)
    assert a.shape == (200, 4)
    # Mach along the wall after 20 iterations from M = 0.8: a stagnation region at the nose, ~0.8 elsewhere
    assert 0.0 <= a[:, 2].min() < 0.3 and a[:, 2].max() < 1.6 and 0.6 < np.median(a[:, 2]) < 1.0
    assert np.isfinite(a).all()
