"""Cost of dfr2d_capture_edge_values (the EdgeQValues store kept for the gradient plot fields) on the C5 step, and the
time of one dfr2d_gradient_field read-back.  usage: python tools/capture_cost.py [nx] [order]  (needs a GPU)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from gocfd_b200 import lib  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    c = bench.build_case(nx, nx, n)
    dev = lib.Dfr2d(c.problem)
    dev.set_state(c.Q)
    out = {"triangles": c.problem.K, "N": n}

    def timed(steps=10):
        dev.step(3)
        t = time.perf_counter()
        dev.step(steps)                 # dfr2d_step synchronises when it reads `info` back
        return (time.perf_counter() - t) / steps * 1e3

    out["ms_per_step_capture_off"] = [timed(), timed()]
    dev.capture_edge_values(True)
    out["ms_per_step_capture_on"] = [timed(), timed()]
    dev.gradient_field(200)
    t = time.perf_counter()
    dev.gradient_field(301)
    out["gradient_field_ms"] = (time.perf_counter() - t) * 1e3
    out["gradient_field_bytes"] = 8 * c.problem.NpFlux * c.problem.K
    dev.capture_edge_values(False)
    out["ms_per_step_capture_off_again"] = [timed()]
    dev.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
