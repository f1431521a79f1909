"""Config C4 as BASELINE.json words it: NACA 0012, M = 0.8, alpha = 2, P = 2 on naca_12/mesh/nacaAirfoil-base.su2 (K = 10,697),
wall / far-field boundaries, Roe flux, local time stepping, PerssonC0 limiter (Kappa 4.5 as in naca_12/input-wall.yaml), run
towards the steady state with the residual history recorded.

  python tools/c4_convergence.py --impl oracle --iterations 3000 --out tests/golden/c4_residual_history.json   (CPU, C oracle)
  python tools/c4_convergence.py --impl device --iterations 30000 --out gpurun_out/c4_device.json                (B200)

The device run compares itself with the golden oracle history at the iterations both hold."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gocfd_b200.host.euler2d import Euler  # noqa: E402
from gocfd_b200.host.input_parameters import InputParameters2D  # noqa: E402

MESH = os.path.join(ROOT, "tests", "golden", "meshes", "nacaAirfoil-base.su2")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "device"], required=True)
    ap.add_argument("--iterations", type=int, default=3000)
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--cfl", type=float, default=2.0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--golden", default=os.path.join(ROOT, "tests", "golden", "c4_residual_history.json"))
    a = ap.parse_args()
    ip = InputParameters2D(Title="C4", CFL=a.cfl, FluxType="Roe", InitType="Freestream", PolynomialOrder=2, FinalTime=1.0e9,
                           Minf=0.8, Gamma=1.4, Alpha=2.0, LocalTimeStepping=True, MaxIterations=a.iterations,
                           Limiter="PerssonC0", Kappa=4.5)
    c = Euler(ip, MESH)
    if a.impl == "oracle":
        from oracle.c_oracle import COracleSolver
        s = COracleSolver(c.problem)
    else:
        from gocfd_b200 import lib
        s = lib.Dfr2d(c.problem)
    s.set_state(c.Q)
    hist = []
    t0 = time.perf_counter()
    done = 0
    while done < a.iterations:
        info = s.step(a.every)
        done = info["steps"]
        r = [float(x) for x in s.residual()]
        q = s.get_state()
        hist.append({"it": done, "residual": r, "l2_rho": float(np.linalg.norm(q[0])), "l2_e": float(np.linalg.norm(q[3]))})
        if not np.isfinite(r).all():
            break
    wall = time.perf_counter() - t0
    out = {"impl": a.impl, "config": "C4: nacaAirfoil-base.su2 K=%d, N=2, Roe, local dt, CFL %g, M 0.8, alpha 2, PerssonC0 Kappa 4.5"
           % (c.problem.K, a.cfl), "iterations": done, "wall_s": wall, "ms_per_step": 1e3 * wall / max(done, 1), "history": hist}
    if a.impl == "device" and os.path.exists(a.golden):
        g = {h["it"]: h for h in json.load(open(a.golden))["history"]}
        worst_r, worst_q, n = 0.0, 0.0, 0
        for h in hist:
            if h["it"] in g:
                ra, rb = np.array(h["residual"]), np.array(g[h["it"]]["residual"])
                worst_r = max(worst_r, float(np.max(np.abs(ra - rb) / np.maximum(np.abs(rb), 1e-300))))
                worst_q = max(worst_q, abs(h["l2_rho"] - g[h["it"]]["l2_rho"]) / g[h["it"]]["l2_rho"])
                n += 1
        out["vs_oracle"] = {"checkpoints": n, "max_rel_residual_diff": worst_r, "max_rel_l2_rho_diff": worst_q}
    with open(a.out, "w") as f:
        json.dump(out, f)
    first, last = hist[0], hist[-1]
    print("its %d  %.2f ms/step  residual max %.3e -> %.3e" % (done, out["ms_per_step"], max(first["residual"]), max(last["residual"])),
          out.get("vs_oracle", ""))


if __name__ == "__main__":
    main()
