"""Partition-local set-up (dfr2d_create_window): a partition built from a WINDOW of the mesh -- its own
element range plus the edge / vertex ring -- must be the partition built from the global problem: identical integer tables
on the CPU (dfr2d_plan_create_window), bitwise identical runs on the device."""
import numpy as np
import pytest

from conftest import mesh_path
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh
from gocfd_b200.host.window import structured_window_case, window_problem, window_range

TABLES = ("kL", "kR", "meta", "etoe", "send_counts", "recv_counts", "ghost_global", "send_elem", "send_row0", "recv_col",
          "recv_row0", "vertex_counts", "vertex_ids")
SIZES = ("k0", "k1", "G", "Kp", "NE", "NEp", "n_cut", "NBP")


def _same_plan(a, b):
    for name in SIZES:
        assert getattr(a, name) == getattr(b, name), name
    for name in TABLES:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def _ip(diss, **kw):
    base = dict(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=2, FinalTime=50.0, MaxIterations=1000, Gamma=1.4,
                Minf=0.1)
    if diss:
        base.update(Limiter="persson c0", Kappa=5.0)
    base.update(kw)
    return InputParameters2D(**base)


@pytest.mark.parametrize("diss", [False, True])
@pytest.mark.parametrize("n_parts", [2, 3, 5])
def test_window_plans_equal_global_plans_on_the_strip_mesh(n_parts, diss):
    """Both ways of getting a window -- cut out of the global Problem, or built from the rows it needs only (what bench.py
    does at 8M elements) -- give the partition plan of the global problem, bit for bit, and the same metrics and state."""
    from gocfd_b200 import lib
    ip = _ip(diss)
    c = Euler(ip, structured_tri_mesh(14, 19))
    for part in range(n_parts):
        g = lib.Plan(c.problem, n_parts, part)
        pw, win = window_problem(c.problem, n_parts, part)
        assert pw.K < c.problem.K and win[0] == c.problem.K
        _same_plan(g, lib.Plan(pw, n_parts, part, window=win))
        cw, win2 = structured_window_case(lambda m: Euler(ip, m), 14, 19, n_parts, part, -10.0, 10.0, -10.0, 10.0)
        _same_plan(g, lib.Plan(cw.problem, n_parts, part, window=win2))
        o = win2[1]
        assert np.array_equal(cw.Q, c.Q[:, :, o:o + cw.problem.K])
        for name in ("Jdet", "EdgeLenMax"):
            assert np.array_equal(getattr(cw.problem, name), getattr(c.problem, name)[o:o + cw.problem.K]), name
        assert np.array_equal(np.asarray(cw.problem.Jinv).reshape(-1, 4), np.asarray(c.problem.Jinv).reshape(-1, 4)[o:o + cw.problem.K])


def test_window_of_an_rcm_ordered_unstructured_mesh():
    """NACA0012 after dfr2d_rcm_order: the numbering is spatially compact, so a partition's ring is a narrow band around
    its own range (on the mesh generator's numbering the window would be the whole mesh)."""
    from gocfd_b200 import lib
    from gocfd_b200.host import readfiles as rf
    mesh = rf.read_mesh(mesh_path("mesh_NACA0012_inv.su2"))
    ip = _ip(True, InitType="Freestream", Minf=0.8, Alpha=2.0, LocalTimeStepping=True)
    c0 = Euler(ip, mesh)
    c1 = Euler(ip, rf.renumber_elements(mesh, lib.rcm_order(c0.problem)))
    for part in range(8):
        w0, w1 = window_range(c1.problem, 8, part)
        assert w1 - w0 < 0.35 * c1.problem.K                # own range is 1/8 of the mesh
        pw, win = window_problem(c1.problem, 8, part)
        _same_plan(lib.Plan(c1.problem, 8, part), lib.Plan(pw, 8, part, window=win))
    u0, u1 = window_range(c0.problem, 8, 3)              # mesh generator's numbering: the ring is scattered far and wide
    r0, r1 = window_range(c1.problem, 8, 3)
    assert u1 - u0 > 2 * (r1 - r0)


def test_window_that_misses_the_partition_is_rejected():
    from gocfd_b200 import lib
    c = Euler(_ip(False), structured_tri_mesh(8, 8))
    pw, win = window_problem(c.problem, 4, 1)
    with pytest.raises(lib.Dfr2dError):
        lib.Plan(pw, 4, 3, window=win)
    with pytest.raises(lib.Dfr2dError):
        lib.Plan(pw, 4, 1, window=(win[0], c.problem.K))    # offset + window size beyond the mesh


@pytest.mark.gpu
@pytest.mark.parametrize("diss", [False, True])
def test_windowed_partitions_run_bitwise_like_the_global_problem(diss):
    """3 partitions, each created from its own window (dfr2d_create_window), stepped with dfr2d_multi_step: the state equals
    the single-partition run of the global problem bit for bit; plot / per-element fields come back in window columns."""
    from gocfd_b200 import lib
    ip = _ip(diss, PolynomialOrder=2 if diss else 3)
    c = Euler(ip, structured_tri_mesh(16, 13))
    if diss:
        c.Q[0] *= 1.0 + 0.1 * np.sign(np.sin(7.0 * c.DFR.solution_xy()[0]))      # jumps inside elements: the sensor fires
    one = lib.Dfr2d(c.problem)
    one.set_state(c.Q)
    one.capture_edge_values(True)
    a = one.step(4)
    want = one.get_state()
    want_grad = {pf: one.gradient_field(pf) for pf in (201, 303)}
    devs, wins = [], []
    for part in range(3):
        pw, win = window_problem(c.problem, 3, part)
        d = lib.Dfr2d(pw, n_parts=3, part=part, window=win)
        d.set_state(np.ascontiguousarray(c.Q[:, :, win[1]:win[1] + pw.K]))
        d.capture_edge_values(True)                       # FaceNorm in window columns, like every other problem array
        devs.append(d)
        wins.append((win[1], pw.K))
    b = lib.multi_step(devs, 4)
    assert a["steps"] == b["steps"] and a["time"] == b["time"]
    got = np.zeros_like(c.Q)
    for d, (off, kw) in zip(devs, wins):
        q = d.get_state()
        k0, k1 = d.partition_range()
        got[:, :, k0:k1] = q[:, :, k0 - off:k1 - off]
        dt_w = d.get_field(0)
        np.testing.assert_array_equal(dt_w[k0 - off:k1 - off], one.get_field(0)[k0:k1])
        for pf, g in want_grad.items():                   # gradient plot fields: the neighbour side reads captured ghost columns
            np.testing.assert_array_equal(d.gradient_field(pf)[:, k0 - off:k1 - off], g[:, k0:k1])
    assert np.array_equal(got, want)
    for d in devs + [one]:
        d.close()
