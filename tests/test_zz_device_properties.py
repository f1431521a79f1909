"""Size-independent properties of full device runs (no oracle).  Needs a GPU.  Kept in its own module, collected last:
these runs were added at the very end of round 1 and have only been through the CPU restatement so far."""
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
