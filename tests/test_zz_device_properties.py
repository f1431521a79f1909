"""Size-independent properties of full device runs (no oracle).  Needs a GPU.  Kept in its own module, collected last:
these runs were added at the very end of round 1 and have only been through the CPU restatement so far."""
import os

import numpy as np
import pytest

from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh

pytestmark = pytest.mark.gpu


def make(ip_kw, mesh):
    base = dict(CFL=1.0, FluxType="Roe", InitType="Freestream", Minf=0.8, Gamma=1.4, Alpha=1.25,
                FinalTime=100.0, MaxIterations=1000)
    base.update(ip_kw)
    return Euler(InputParameters2D(**base), mesh)


@pytest.mark.parametrize("n,flux,coarse,fine,cfl,min_order,max_fine_error", [
    (2, "Roe", 24, 48, 0.5, 2.5, 8e-4),
    (4, "Lax", 24, 48, 0.25, 4.3, 4e-5),
])
def test_device_converges_to_the_exact_vortex(n, flux, coarse, fine, cfl, min_order, max_fine_error):
    """Size-independent property, no oracle: the device run to t = 0.5 converges to the analytic isentropic vortex at
    high order (the CPU restatement measures 2.75 at N=2 Roe and 4.78 at N=4 Lax on these meshes, 1.6e-5 at N=4 fine;
    tests/test_physics_pins.py)."""
    from gocfd_b200 import lib

    def error(nx):
        c = make(dict(PolynomialOrder=n, FluxType=flux, InitType="IVortex", CFL=cfl, FinalTime=0.5, MaxIterations=10 ** 6,
                      Minf=0.1), structured_tri_mesh(nx, nx, tag="wall"))
        dev = lib.Dfr2d(c.problem)
        dev.set_state(c.Q)
        info = dev.step(100)
        while not info["finished"]:          # (a single huge count would enqueue no-op stages long after FinalTime)
            info = dev.step(100)
        assert info["time"] == pytest.approx(0.5, abs=1e-12)
        q = dev.get_state()
        dev.close()
        x, y = c.DFR.solution_xy()
        exact = np.zeros_like(q)
        for i in range(x.shape[0]):
            for k in range(x.shape[1]):
                exact[:, i, k] = c.AnalyticSolution.get_state_c(info["time"], x[i, k], y[i, k])
        w = c.DFR.Jdet[None, None, :]
        return float(np.sqrt((w * (q - exact) ** 2).sum() / (w * np.ones_like(q)).sum()))

    e1, e2 = error(coarse), error(fine)
    assert e2 < max_fine_error
    assert np.log2(e1 / e2) > min_order


# ---- dfr2d_multi_step: the single-process multi-GPU driver.  Written at the end of round 1 without a GPU at hand, so
# it only runs on request until it has been through one GPU run; here all partitions share one device (peer copies
# degenerate to device copies, the event waits to stream order) and the result must be the single-partition run, bit
# for bit, like the hand-driven protocol of tests/test_gpu_parity.py.

def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("spread", [False, True])
@pytest.mark.parametrize("n_parts,n,diss", [(3, 2, False), (2, 2, True), (4, 4, True)])
def test_multi_step_driver_matches_single_partition(n_parts, n, diss, spread):
    """dfr2d_multi_step (mailbox protocol of csrc/dfr2d_peer.cuh: P2P puts + arrival flags + wave inbox) must reproduce
    the single-partition run bit for bit.  spread=False: all partitions on device 0 (one stream: the issue order
    'all puts before any wait' is what keeps it deadlock free).  spread=True: partition g on device g mod #devices --
    real peer stores over NVLink; needs >= 2 GPUs."""
    from conftest import mesh_path
    from gocfd_b200 import lib
    ndev = _n_devices()
    if spread and ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    if diss:
        c = make(dict(PolynomialOrder=n, InitType="shocktube", CFL=2.0, FinalTime=0.2, Limiter="persson c0", Kappa=5.0),
                 mesh_path("sod-aligned-100pts.su2"))
        x, _ = c.DFR.solution_xy()
        w = 0.5 * (1.0 - np.tanh((x - 0.503) / 0.004))
        c.Q = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w for v in range(4)])
    else:
        c = make(dict(PolynomialOrder=n, InitType="IVortex", CFL=1.0, FinalTime=50.0), structured_tri_mesh(16, 12))
    one = lib.Dfr2d(c.problem)
    one.set_state(c.Q)
    a = one.step(4)
    devs = [lib.Dfr2d(c.problem, n_parts=n_parts, part=r, device=(r % ndev) if spread else 0) for r in range(n_parts)]
    lib.multi_set_state(devs, c.Q)
    lib.multi_step(devs, 1, sync=False)                  # no host synchronisation without info
    prof = lib.multi_step_profile(devs)                  # the profiled step is an ordinary step
    assert prof.shape == (n_parts, 5, len(lib.PROFILE_PHASES)) and (prof >= 0).all() and prof.sum() > 0
    b = lib.multi_step(devs, 2)
    assert a["steps"] == b["steps"] == 4 and a["time"] == b["time"] and a["dt"] == b["dt"]
    q = lib.multi_get_state(devs)
    assert np.array_equal(q, one.get_state())
    # rewind the clock (dfr2d_set_clock) and repeat: the IVortex boundary state depends on rk.Time, so only a run that
    # restarts at t = 0 reproduces the first one
    for d in devs:
        d.set_clock(0.0, 0)
    lib.multi_set_state(devs, c.Q)
    b2 = lib.multi_step(devs, 4)
    assert b2["steps"] == 4 and b2["time"] == a["time"]
    assert np.array_equal(lib.multi_get_state(devs), q)
    for d in devs:
        d.close()
    one.close()


@pytest.mark.parametrize("world,diss", [(2, False), (3, True)])
def test_peer_connected_processes_match_single_partition(world, diss, tmp_path):
    """One PROCESS per partition (the torchrun shape): every process creates its partition, the mailbox descriptions are
    all-gathered (gloo), dfr2d_peer_connect maps the partners' mailboxes through CUDA IPC, then each process simply calls
    dfr2d_step.  No NCCL, no host in the stage loop.  Partitions use device rank mod #devices (on a one-GPU box the
    processes time-slice the device).  Must equal the single-partition run bit for bit."""
    import subprocess
    import sys
    from conftest import ROOT
    from gocfd_b200 import lib
    port = 29600 + (os.getpid() % 300)
    out = str(tmp_path / "state")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "peer_ipc_worker.py"), str(r), str(world),
                               str(int(diss)), str(port), out], cwd=ROOT) for r in range(world)]
    try:
        rcs = [p.wait(timeout=240) for p in procs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert rcs == [0] * world
    from peer_ipc_worker import build
    c = build(diss)
    one = lib.Dfr2d(c.problem)
    one.set_state(c.Q)
    a = one.step(4)
    q = np.zeros_like(c.Q)
    for r in range(world):
        part = np.load("%s_%d.npz" % (out, r))
        q[:, :, int(part["k0"]):int(part["k1"])] = part["q"]
        assert float(part["time"]) == a["time"] and int(part["steps"]) == 4
    assert np.array_equal(q, one.get_state())
    one.close()


@pytest.mark.parametrize("diss_elem", [1, 3, 5])
def test_naca_front_local_dt_active_dissipation_on_device(diss_elem, monkeypatch):
    """Local time stepping with an ACTIVE sensor (tests/test_c_oracle.py::_naca_front_case: ~400 elements above
    sigma = 0.05 in the first steps, DTVisc > 1e-9 in ~500): viscous dt limit and DTVisc carry-over on the device."""
    from test_c_oracle import _naca_front_case
    from gocfd_b200 import lib
    from oracle.euler2d_oracle import OracleSolver
    monkeypatch.setenv("DFR2D_DISS_ELEM_KERNEL", str(diss_elem))
    c = _naca_front_case()
    dev, ora = lib.Dfr2d(c.problem), OracleSolver(c.problem)
    dev.set_state(c.Q)
    ora.set_state(c.Q)
    dev.step(3), ora.step(3)
    assert (ora.DTVisc > 1e-9).sum() > 100
    den = np.linalg.norm(ora.get_state())
    assert np.linalg.norm(dev.get_state() - ora.get_state()) / den < 1e-11
    np.testing.assert_allclose(dev.get_field(0), ora.DT, rtol=1e-10)
    dev.close()


@pytest.mark.parametrize("diss", [False, True])
def test_scattered_partitions_on_device(diss):
    """Elements renumbered at random, 5 partitions: contiguous Split1D ranges are scattered in space, so almost every edge
    is cut and every partition is mostly halo -- the worst case for the ghost columns / halo lists the kernels consume
    (their integer tables are checked bit-exactly on the CPU, tests/test_abi_and_partition.py).  Must equal the
    single-partition device run bitwise and the oracle to 1e-11."""
    from test_abi_and_partition import _shuffled_grid
    from test_gpu_parity import _multi_partition_step
    from gocfd_b200 import lib
    from oracle.euler2d_oracle import OracleSolver
    kw = dict(PolynomialOrder=2, InitType="IVortex", CFL=1.0, FinalTime=50.0)
    if diss:
        kw.update(Limiter="PerssonC0", Kappa=5.0)
    c = make(kw, _shuffled_grid(5))
    if diss:        # sharpen the vortex core so that the sensor fires somewhere
        c.Q[0] *= 1.0 + 0.3 * np.sign(np.sin(7.0 * c.DFR.solution_xy()[0]))
    devs = [lib.Dfr2d(c.problem, n_parts=5, part=r) for r in range(5)]
    for d in devs:
        d.set_state(c.Q)
    _multi_partition_step(devs, 3)
    q = np.zeros_like(c.Q)
    for d in devs:
        d.get_state(q)
    one = lib.Dfr2d(c.problem)
    one.set_state(c.Q)
    one.step(3)
    assert np.array_equal(q, one.get_state())
    ora = OracleSolver(c.problem)
    ora.set_state(c.Q)
    ora.step(3)
    if diss:
        assert ora.SigmaScalar.max() > 0.02
    assert np.linalg.norm(q - ora.get_state()) / np.linalg.norm(ora.get_state()) < 1e-11
    for d in devs + [one]:
        d.close()


@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("flux", ["Roe", "Lax", "roe-er", "average"])
def test_dissipation_flux_and_dt_matrix_on_device(flux, local):
    """The PerssonC0 path with every numerical flux and both dt modes on the Sod mesh at N=2 (the GPU dissipation tests
    of test_gpu_parity.py all use Roe + global dt).  The tube starts at rest, which puts Roe-ER on its u = v = 0 branch
    (1/|V| is infinite and must not leak through the select, fluxes.go:440-450)."""
    from conftest import mesh_path
    from gocfd_b200 import lib
    from oracle.euler2d_oracle import OracleSolver
    c = make(dict(PolynomialOrder=2, InitType="shocktube", CFL=1.0, FinalTime=0.2, Limiter="persson c0", Kappa=5.0,
                  FluxType=flux, LocalTimeStepping=local, MaxIterations=1000), mesh_path("sod-aligned-100pts.su2"))
    x, _ = c.DFR.solution_xy()
    w = 0.5 * (1.0 - np.tanh((x - 0.503) / 0.002))
    q = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w for v in range(4)])
    dev, ora = lib.Dfr2d(c.problem), OracleSolver(c.problem)
    dev.set_state(q)
    ora.set_state(q)
    dev.step(2), ora.step(2)
    assert ora.SigmaScalar.max() > 0.05 and np.isfinite(ora.get_state()).all()
    assert np.linalg.norm(dev.get_state() - ora.get_state()) / np.linalg.norm(ora.get_state()) < 1e-11
    dev.close()
