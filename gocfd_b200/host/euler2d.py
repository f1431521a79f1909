"""Host-side mirror of the reference's `Euler2D` set-up: what the Go host does before (and
around) the time loop.  It builds the flat problem description that crosses the C ABI
(include/dfr2d.h) and drives the device library through `gocfd_b200.lib`.

Nothing here is on the hot path: operator/geometry construction and the initial condition
run once; `Euler.Solve` only calls `dfr2d_step` and prints the reference's progress table.

Reference: model_problems/Euler2D/euler.go:61-136 (NewEuler), :138-220 (Solve), :728-849
(InitializeSolution, CheckIfFinished, PrintUpdate); fluids.go:237-273 (FreeStream);
fluxes.go:18-51, initialization.go:17-83, filter.go:8-48 (enums);
isentropic_vortex/analytic_vortex.go:31-82; utils/parallel_utils.go:86-192 (PartitionMap);
parallelism.go:100-120 (SetParallelDegree).
"""
import math
import time

import numpy as np

from .dg2d.dfr2d import DFR2D
from . import readfiles as rf
from .input_parameters import InputParameters2D

FLUX_Average, FLUX_LaxFriedrichs, FLUX_Roe, FLUX_RoeER = range(4)
FLUX_NAMES = {"average": FLUX_Average, "lax": FLUX_LaxFriedrichs, "roe": FLUX_Roe, "roe-er": FLUX_RoeER}
FREESTREAM, IVORTEX, SHOCKTUBE = range(3)
INIT_NAMES = {"freestream": FREESTREAM, "ivortex": IVORTEX, "shocktube": SHOCKTUBE}
LIMITER_None, LIMITER_PerssonC0 = range(2)
LIMITER_NAMES = {"perssonc0": LIMITER_PerssonC0, "persson c0": LIMITER_PerssonC0}


def new_flux_type(label):
    key = label.lower()
    if key not in FLUX_NAMES:
        raise ValueError("unable to use flux named %s" % key)
    return FLUX_NAMES[key]


def new_init_type(label):
    if len(label) == 0:
        raise ValueError("empty init type, must be one of %s" % sorted(INIT_NAMES))
    key = label.lower()
    if key not in INIT_NAMES:
        raise ValueError("unable to use init type named %s" % key)
    return INIT_NAMES[key]


def new_limiter_type(label):
    if len(label) == 0:
        return LIMITER_None
    key = label.strip().lower()
    if key not in LIMITER_NAMES:
        raise ValueError("unable to use limiter named [%s]" % key)
    return LIMITER_NAMES[key]


class FreeStream:
    def __init__(self, gamma, qinf, minf=0.0, alpha=0.0):
        self.Gamma = gamma
        self.Qinf = np.array(qinf, dtype=np.float64)
        self.Alpha = alpha
        self.Minf = minf
        rho, rho_u, rho_v, e = (float(x) for x in self.Qinf)
        oorho = 1.0 / rho
        u, v = rho_u * oorho, rho_v * oorho
        u2 = u * u + v * v
        q = 0.5 * rho * u2
        p = (gamma - 1.0) * (e - q)
        self.Pinf = p
        self.QQinf = q
        self.Cinf = math.sqrt(abs(gamma * p * oorho))

    @classmethod
    def from_mach(cls, minf, gamma, alpha):
        ooggm1 = 1.0 / (gamma * (gamma - 1.0))
        uinf = minf * math.cos(alpha * math.pi / 180.0)
        vinf = minf * math.sin(alpha * math.pi / 180.0)
        return cls(gamma, [1.0, uinf, vinf, ooggm1 + 0.5 * minf * minf], minf, alpha)

    def as_array(self):
        """[Gamma, Qinf0..3, Pinf, QQinf, Cinf, Alpha, Minf] -- layout of dfr2d_freestream."""
        return np.array([self.Gamma, *self.Qinf, self.Pinf, self.QQinf, self.Cinf, self.Alpha, self.Minf])


class PartitionMap:
    """Contiguous ranges over [0, max_index), sizes differing by at most one."""

    def __init__(self, parallel_degree, max_index):
        self.ParallelDegree = parallel_degree
        self.MaxIndex = max_index
        self.Partitions = [self.split_1d(n) for n in range(parallel_degree)]

    def split_1d(self, n):
        npart = self.MaxIndex // self.ParallelDegree
        rem = self.MaxIndex % self.ParallelDegree
        start_add, end_add = 0, 0
        if rem != 0:
            if n + 1 > rem:
                start_add, end_add = rem, 0
            else:
                start_add, end_add = n, 1
        lo = n * npart + start_add
        return (lo, lo + npart + end_add)

    def get_bucket_range(self, bn):
        return self.Partitions[bn]

    def get_bucket_dimension(self, bn):
        if bn == -1:
            return self.MaxIndex
        lo, hi = self.Partitions[bn]
        return hi - lo

    def get_bucket(self, k):
        bn = int(float(self.ParallelDegree * k) / float(self.MaxIndex))
        while not (self.Partitions[bn][0] <= k < self.Partitions[bn][1]):
            bn += -1 if self.Partitions[bn][0] > k else 1
            if bn == -1 or bn == self.ParallelDegree:
                return -1, 0, 0
        return bn, self.Partitions[bn][0], self.Partitions[bn][1]

    def get_local_k(self, base_k):
        bn, lo, hi = self.get_bucket(base_k)
        return base_k - lo, hi - lo, bn

    def get_global_k(self, k_local, bn):
        return k_local if bn == -1 else self.Partitions[bn][0] + k_local


class IVortex:
    def __init__(self, beta=5.0, x0=5.0, y0=0.0, gamma=1.4, ufs=1.0):
        self.Beta, self.X0, self.Y0, self.Gamma, self.Ufs = beta, x0, y0, gamma, ufs

    def get_state_c(self, t, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        oo2pi = 0.5 * (1.0 / math.pi)
        gm1 = self.Gamma - 1.0
        oogm1 = 1.0 / gm1
        fac = 16.0 * self.Gamma * (math.pi * math.pi)
        beta, beta2 = self.Beta, self.Beta * self.Beta
        u0, v0 = self.Ufs, 0.0
        xmut, ymvt = x - u0 * t, y - v0 * t
        r2 = (xmut - self.X0) * (xmut - self.X0) + (ymvt - self.Y0) * (ymvt - self.Y0)
        ex1r = np.exp(1.0 - r2)
        tv1 = 1.0 - (gm1 * beta2 * np.exp(2.0 * (1.0 - r2)) / fac)
        u = u0 - beta * ex1r * (ymvt - self.Y0) * oo2pi
        v = v0 + beta * ex1r * (xmut - self.X0) * oo2pi
        rho = np.power(tv1, oogm1)
        p = np.power(rho, self.Gamma)
        q = 0.5 * rho * (u * u + v * v)
        return rho, rho * u, rho * v, p * (1.0 / (self.Gamma - 1.0)) + q

    def as_array(self):
        return np.array([self.Beta, self.X0, self.Y0, self.Gamma, self.Ufs])


class Problem:
    """Flat, library-facing description of one solver instance (fields of `dfr2d_problem`)."""

    def __init__(self):
        pass


class Euler:
    """NewEuler + Solve.  `backend` is a gocfd_b200.lib.Dfr2d-like object factory; the tests also
    pass the oracle here so both sides consume the identical Problem."""

    def __init__(self, ip: InputParameters2D, mesh, proc_limit=1, verbose=False):
        self.ip = ip
        self.CFL = ip.CFL
        self.FinalTime = ip.FinalTime
        self.FluxCalcAlgo = new_flux_type(ip.FluxType)
        self.Case = new_init_type(ip.InitType)
        self.LocalTimeStepping = ip.LocalTimeStepping
        self.MaxIterations = ip.MaxIterations
        self.FSFar = FreeStream.from_mach(ip.Minf, ip.Gamma, ip.Alpha)
        self.FSIn = None
        self.FSOut = None
        self.Kappa = ip.Kappa
        self.AnalyticSolution = None
        if isinstance(mesh, str):
            mesh = rf.read_mesh(mesh)
        self.DFR = DFR2D(ip.PolynomialOrder, mesh)
        self.ParallelDegree = max(1, proc_limit)
        if self.ParallelDegree > self.DFR.K:
            self.ParallelDegree = 1
        self.Partitions = PartitionMap(self.ParallelDegree, self.DFR.K)
        self._initialize_solution(verbose)
        lt = new_limiter_type(ip.Limiter)
        self.Dissipation = (lt == LIMITER_PerssonC0 and self.DFR.N != 0)
        self.problem = self._build_problem()

    # ---- initial conditions (euler.go:728-794) -----------------------------------------
    def _initialize_solution(self, verbose):
        dfr = self.DFR
        el = dfr.SolutionElement
        k = dfr.K
        q = np.empty((4, el.Np, k))
        if self.Case == FREESTREAM:
            for n in range(4):
                q[n] = self.FSFar.Qinf[n]
        elif self.Case == SHOCKTUBE:
            gamma = 1.4
            self.FSIn = FreeStream(gamma, [1.0, 0.0, 0.0, 1.0 / (gamma - 1.0)])
            self.FSOut = FreeStream(gamma, [0.125, 0.0, 0.0, 0.1 / (gamma - 1.0)])
            for i in range(el.Np):
                x, _ = dfr.local_coords(el.R[i:i + 1], el.S[i:i + 1])
                left = x[0] < 0.5
                for n in range(4):
                    q[n, i] = np.where(left, self.FSIn.Qinf[n], self.FSOut.Qinf[n])
        elif self.Case == IVORTEX:
            self.FSFar = FreeStream(1.4, [1.0, 1.0, 0.0, 3.0])
            self.AnalyticSolution = IVortex(5.0, 5.0, 0.0, 1.4)
            for i in range(el.Np):
                x, y = dfr.local_coords(el.R[i:i + 1], el.S[i:i + 1])
                st = self.AnalyticSolution.get_state_c(0.0, x[0], y[0])
                for n in range(4):
                    q[n, i] = st[n]
            t = dfr.Tris
            wall = t.bcType == rf.BC_Wall
            t.bcType[wall] = rf.BC_IVortex
            if verbose:
                print("\tReplaced %d Wall boundary conditions with analytic BC_IVortex" % int(wall.sum()))
        else:
            raise ValueError("unknown case type")
        self.Q = q

    # ---- flatten for the C ABI ---------------------------------
    def _build_problem(self):
        dfr = self.DFR
        rt = dfr.FluxElement
        el = dfr.SolutionElement
        t = dfr.Tris
        p = Problem()
        p.N, p.K, p.NV, p.NE = dfr.N, dfr.K, len(dfr.VX), t.NE
        p.NpInt, p.NpEdge, p.NpFlux = el.Np, rt.NpEdge, rt.Np
        c = np.ascontiguousarray
        p.FluxEdgeInterp = c(dfr.FluxEdgeInterp)
        p.DivInt = c(rt.DivInt)
        p.Div = c(rt.Div)
        p.V, p.Vinv = c(el.JB2D.V), c(el.JB2D.Vinv)
        sf = dfr.shock_finder()
        p.MassMatrix, p.D, p.P = c(sf.MassMatrix), c(sf.D), c(sf.P)
        p.ModeFilter = c(sf.ModeFilter)
        p.Bary = c(dfr.barycentric_coords())
        p.Jdet, p.Jinv = c(dfr.Jdet), c(dfr.Jinv)
        p.FaceNormX, p.FaceNormY = c(dfr.FaceNorm[0]), c(dfr.FaceNorm[1])
        p.IInII = c(dfr.IInII)
        p.EdgeLenMax = c(dfr.EdgeLenMax)
        p.EToV = c(dfr.EToV.astype(np.int32))
        p.edge_kL, p.edge_kR = c(t.kL), c(t.kR)
        p.edge_numL, p.edge_numR = c(t.edgeNumL), c(t.edgeNumR)
        p.edge_nconn, p.edge_bc = c(t.nConn), c(t.bcType)
        p.edge_len = c(dfr.edge_length())
        p.EtoEdge = c(t.EtoEdge)
        # coordinates of the edge points of boundary edges (IVortex BC reads FluxX/FluxY there)
        bidx = np.flatnonzero(t.nConn == 1).astype(np.int32)
        ne = rt.NpEdge
        off = 2 * rt.NpInt
        bx = np.empty((len(bidx), ne))
        by = np.empty((len(bidx), ne))
        if len(bidx):
            kb, eb = t.kL[bidx], t.edgeNumL[bidx]
            for e in range(3):
                sel = eb == e
                if not sel.any():
                    continue
                rows = slice(off + e * ne, off + (e + 1) * ne)
                x, y = dfr.local_coords(rt.R[rows], rt.S[rows], kb[sel])
                bx[sel], by[sel] = x.T, y.T
        p.NBP, p.bp_edge, p.bp_x, p.bp_y = len(bidx), bidx, c(bx), c(by)
        p.FSFar = self.FSFar
        p.FSIn = self.FSIn if self.FSIn is not None else self.FSFar
        p.FSOut = self.FSOut if self.FSOut is not None else self.FSFar
        p.Gamma = self.FSFar.Gamma
        p.FluxType, p.Case = self.FluxCalcAlgo, self.Case
        p.LocalTimeStepping = bool(self.LocalTimeStepping)
        p.MaxIterations = int(self.MaxIterations)
        p.Dissipation = bool(self.Dissipation)
        p.CFL, p.FinalTime, p.Kappa = float(self.CFL), float(self.FinalTime), float(self.Kappa)
        p.Vortex = self.AnalyticSolution if self.AnalyticSolution is not None else IVortex()
        return p

    # ---- time loop (euler.go:138-220) -------------------------------------------------
    def Solve(self, solver, print_every=100, out=print, output_dir=None, mesh_file=""):
        """`solver` implements set_state/step/residual/get_state (the device library or the oracle).  With `output_dir`
        the files of `OutputFinal` (euler.go:219-318) are written after the last step, as the reference does."""
        solver.set_state(self.Q)
        out(self.print_initialization())
        steps, finished = 0, False
        elapsed = 0.0
        while not finished:
            t0 = time.perf_counter()
            chunk = 1 if steps == 0 else print_every - (steps % print_every)
            info = solver.step(chunk)
            elapsed += time.perf_counter() - t0
            steps = info["steps"]
            finished = info["finished"]
            if finished or steps == 1 or steps % print_every == 0:
                out(self.print_update(info["time"], info["dt"], steps, solver.residual()))
        self.Q = solver.get_state()
        rate = elapsed * 1e6 / float(self.DFR.K * max(steps, 1))
        out("\nRate of execution = %8.5f us/(element*iteration) over %d iterations" % (rate, steps))
        if output_dir is not None:
            from host_standin.output_final import output_final     # post-processing stays host-side (stand-in)
            output_final(self, self.Q, mesh_file=mesh_file, outdir=output_dir, out=out)
        return steps, elapsed

    def check_if_finished(self, t, steps):
        return t >= self.FinalTime or steps >= self.MaxIterations

    def print_initialization(self):
        if self.LocalTimeStepping:
            s = "Solving until Max Iterations = %d\n    iter                " % self.MaxIterations
        else:
            s = "Solving until finaltime = %8.5f\n    iter    time      dt" % self.FinalTime
        return s + "       Res0       Res1       Res2       Res3         L1         L2"

    def print_update(self, t, dt, steps, max_r):
        if self.LocalTimeStepping:
            s = "%10d              " % steps
        else:
            s = "%8d%8.5f%8.5f" % (steps, t, dt)
        l1, l2 = 0.0, 0.0
        for n in range(4):
            m = max(0.0, float(max_r[n]))     # euler.go:823-829: maxR starts at 0
            s += "%11.4e" % m
            l1 = max(l1, m)
            l2 += m * m
        return s + "%11.4e%11.4e" % (l1, math.sqrt(l2) / 4.0)


def new_euler(ip, mesh_file, proc_limit=1, verbose=False):
    return Euler(ip, mesh_file, proc_limit, verbose)
