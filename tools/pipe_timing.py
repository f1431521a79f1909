"""Per-phase clock64 breakdown of k_elem_pipe (build with -DDFR2D_PIPE_TIMING)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gocfd_b200 import lib
lib.LIB_PATH = os.path.join(os.path.dirname(lib.LIB_PATH), "libdfr2d_timing.so")
import bench
c = bench.build_case(1000, 1000, 4)
d = lib.Dfr2d(c.problem)
d.set_state(c.Q)
d.step(2)
out = (ctypes.c_ulonglong * 8)()
d.lib.dfr2d_debug_pipe_clocks(out, 1)
d.step(2)
d.lib.dfr2d_debug_pipe_clocks(out, 0)
v = np.array(list(out), dtype=float)
names = ["top barrier wait", "cp.async issue", "regs->smem", "issue loads t+1", "barrier 2", "flux + cp.wait + barrier 3", "DMMA1 + epilogue", "DMMA2 + qface stores"]
ntile_iters = 2 * 5 * (c.problem.K / 32)       # tile iterations summed over CTAs
for n, x in zip(names, v):
    print("%-28s %8.0f cycles/tile  %5.1f%%" % (n, x / ntile_iters, 100 * x / v.sum()))
print("total per tile iteration: %.0f cycles" % (v.sum() / ntile_iters))
