"""Sod shock-tube post-processing of `Euler.OutputFinal` -- host side, runs once after the
last step on the state read back with dfr2d_get_state.

Mirrors, with the reference's names and quirks:
  * `SOD_Exact` (model_problems/Euler1D/sod_shock_tube/analytic_sod.go:21-185): exact Riemann solution, the secant
    root finder `fzero` (start pi, previous start/2, tolerance 1e-7 on |f|) and the 20 sample abscissae of `Get`.
  * `SODShockTube` (model_problems/Euler2D/sod_shock_tube/shock_tube.go:13-142): nPts = 4K/5 centre-line sample points,
    element search by barycentric test in element order, and the interpolation row
    `JB2D.GetInterpMatrix(r, s)` built from the barycentric weights themselves -- the reference passes the [0,1]
    weights (r along v2-v0, s along v1-v0), not [-1,1] reference coordinates, to the basis; reproduced as written.
  * the two files of `OutputFinal` (euler.go:221-250): shocktube.dat / shocktube_analytic.dat, "%.8f" columns.
"""
import math

import numpy as np


class State:
    def __init__(self, rho, p, u, gamma):
        self.rho, self.p, self.u, self.gamma = rho, p, u, gamma

    def C(self):
        return math.sqrt(self.gamma * self.p / self.rho)


def sod_func(P):
    """Pressure function whose root is the post-shock pressure (analytic_sod.go:175-185)."""
    rho_l, P_l = 1.0, 1.0
    rho_r, P_r = 0.125, 0.1
    gamma = 1.4
    mu = math.sqrt((gamma - 1) / (gamma + 1))
    mu2 = mu * mu
    return ((P - P_r) * math.sqrt((1 - mu2) / (rho_r * (P + mu2 * P_r)))
            - (math.pow(P_l, (gamma - 1) / (2 * gamma)) - math.pow(P, (gamma - 1) / (2 * gamma)))
            * math.sqrt(((1 - mu2 * mu2) * math.pow(P_l, 1 / gamma)) / (mu2 * mu2 * rho_l)))


def fzero(f, start):
    """Secant iteration of analytic_sod.go:156-173 (|start - f*deriv| keeps the iterate positive)."""
    tol = 0.0000001
    start_old = start / 2
    res = f(start_old)
    while abs(res) > tol:
        res_new = f(start)
        deriv = (start - start_old) / (res_new - res)
        start_new = abs(start - res_new * deriv)
        start_old = start
        start = start_new
        res = res_new
    return start


class SODExact:
    """NewSOD(t) (analytic_sod.go:31-68)."""

    def __init__(self, t):
        gamma = 1.4
        self.gamma = gamma
        self.x_min, self.x_max = 0.0, 1.0
        self.l_s = State(1.0, 1.0, 0.0, gamma)
        self.r_s = State(0.125, 0.1, 0.0, gamma)
        self.calc(t)

    def calc(self, t):
        gamma = self.gamma
        mu = math.sqrt((gamma - 1) / (gamma + 1))
        mu2 = mu * mu
        l_s, r_s = self.l_s, self.r_s
        p_post = fzero(sod_func, math.pi)
        self.t = t
        self.rho_middle = l_s.rho * math.pow(p_post / l_s.p, 1.0 / gamma)
        self.x0 = 0.5 * (self.x_max + self.x_min)
        self.post_s = State(
            r_s.rho * ((p_post / r_s.p) + mu2) / (1 + mu2 * (p_post / r_s.p)),
            p_post,
            r_s.u + (p_post - r_s.p) / math.sqrt(0.5 * r_s.rho * ((gamma + 1) * p_post + (gamma - 1) * r_s.p)),
            gamma)
        ratio = self.post_s.rho / r_s.rho
        v_shock = self.post_s.u * ratio / (ratio - 1.0)
        self.x1 = self.x0 - l_s.C() * t
        self.x3 = self.x0 + self.post_s.u * t
        self.x4 = self.x0 + v_shock * t
        c_2 = l_s.C() - 0.5 * (gamma - 1.0) * self.post_s.u
        self.x2 = self.x0 + t * (self.post_s.u - c_2)

    def getx(self, x):
        """(rho, p, u, e, rhou) at x; the `switch` of analytic_sod.go:79-103 takes the first matching case."""
        gamma = self.gamma
        mu = math.sqrt((gamma - 1) / (gamma + 1))
        mu2 = mu * mu
        l_s, r_s = self.l_s, self.r_s
        x0, x1, x2, x3, x4 = self.x0, self.x1, self.x2, self.x3, self.x4
        rho = p = u = 0.0
        if x < x1:
            rho, p, u = l_s.rho, l_s.p, l_s.u
        elif x1 <= x <= x2:
            c = mu2 * ((x0 - x) / self.t) + (1.0 - mu2) * l_s.C()
            rho = l_s.rho * math.pow(c / l_s.C(), 2 / (gamma - 1))
            p = l_s.p * math.pow(rho / l_s.rho, gamma)
            u = (1.0 - mu2) * ((-(x0 - x) / self.t) + l_s.C())
        elif x2 <= x <= x3:
            rho, p, u = self.rho_middle, self.post_s.p, self.post_s.u
        elif x3 <= x <= x4:
            rho, p, u = self.post_s.rho, self.post_s.p, self.post_s.u
        elif x4 < x:
            rho, p, u = r_s.rho, r_s.p, r_s.u
        e = p / (gamma - 1.0) + 0.5 * u * u * rho
        return rho, p, u, e, rho * u

    def get(self):
        """(X, Rho, P, RhoU, E) on the 20 abscissae of analytic_sod.go:116-128."""
        x1, x2, x3, x4 = self.x1, self.x2, self.x3, self.x4
        tol = 0.0001
        mid = (x2 - x1) / 10.0
        X = [self.x_min, x1 - tol, x1 + tol]
        X += [x1 + m * mid for m in range(1, 10)]
        X += [x1 + 10 * mid - 2.0 * tol, x2 - tol, x2 + tol, x3 - tol, x3 + tol, x4 - tol, x4 + tol, self.x_max]
        vals = [self.getx(x) for x in X]
        rho = [v[0] for v in vals]
        p = [v[1] for v in vals]
        e = [v[3] for v in vals]
        rhou = [v[4] for v in vals]
        return np.array(X), np.array(rho), np.array(p), np.array(rhou), np.array(e)


class SODShockTube:
    """NewSODShockTube(nPts, dfr) (shock_tube.go:27-49)."""

    def __init__(self, n_pts, dfr):
        self.Npts = n_pts
        self.DFR2D = dfr
        xfrac = 1.0 / float(n_pts - 1)
        x = np.arange(n_pts, dtype=np.float64) * xfrac
        x[0] += 0.00001
        x[n_pts - 1] -= 0.00001
        self.XLocations = x
        self.Rho = np.zeros(n_pts)
        self.RhoU = np.zeros(n_pts)
        self.E = np.zeros(n_pts)
        self._calculate_interpolation()

    def get_analytic_solution(self, t):
        x, rho, p, rhou, e = SODExact(t).get()
        return x, rho, p, rhou, e

    def _calculate_interpolation(self):
        dfr = self.DFR2D
        vy = dfr.VY
        ymid = 0.5 * (vy.max() - vy.min()) + vy.min()
        self.ElementNumber = np.zeros(self.Npts, dtype=np.int64)
        self.RS = np.zeros((self.Npts, 2))
        jb2d = dfr.SolutionElement.JB2D
        self.InterpolationMatrix = np.zeros((self.Npts, dfr.SolutionElement.Np))
        # the element search of getUVCoords, vectorised over elements; np.argmax keeps "first element in order"
        etov = dfr.EToV
        ax, ay = dfr.VX[etov[:, 0]], dfr.VY[etov[:, 0]]
        v0x, v0y = dfr.VX[etov[:, 2]] - ax, dfr.VY[etov[:, 2]] - ay          # C - A
        v1x, v1y = dfr.VX[etov[:, 1]] - ax, dfr.VY[etov[:, 1]] - ay          # B - A
        dot00 = v0x * v0x + v0y * v0y
        dot01 = v0x * v1x + v0y * v1y
        dot11 = v1x * v1x + v1y * v1y
        inv_denom = 1.0 / (dot00 * dot11 - dot01 * dot01)
        for i, x in enumerate(self.XLocations):
            v2x, v2y = x - ax, ymid - ay
            dot02 = v0x * v2x + v0y * v2y
            dot12 = v1x * v2x + v1y * v2y
            r = (dot11 * dot02 - dot01 * dot12) * inv_denom
            s = (dot00 * dot12 - dot01 * dot02) * inv_denom
            inside = (r >= 0) & (s >= 0) & ((r + s) <= 1.0)
            if not inside.any():
                raise RuntimeError("unable to find point within elements: [%5.3f,%5.3f]" % (x, ymid))
            k = int(np.argmax(inside))
            self.ElementNumber[i] = k
            self.RS[i] = (r[k], s[k])
            self.InterpolationMatrix[i] = jb2d.interp_matrix(np.array([r[k]]), np.array([s[k]]))[0]

    def interpolate_fields(self, q):
        """InterpolateFields (shock_tube.go:124-142): rho, rho*u and E (not rho*v) at the sample points."""
        k = self.ElementNumber
        im = self.InterpolationMatrix
        self.Rho = np.einsum("ij,ji->i", im, q[0][:, k])
        self.RhoU = np.einsum("ij,ji->i", im, q[1][:, k])
        self.E = np.einsum("ij,ji->i", im, q[3][:, k])


def shocktube_files(st, mesh_file, final_time):
    """Contents of shocktube.dat and shocktube_analytic.dat as written by OutputFinal (euler.go:230-250)."""
    lines = ["Meshfile: %s\n" % mesh_file, "X\tRho\tRhoU\tE\n"]
    for i, x in enumerate(st.XLocations):
        lines.append("%.8f\t%.8f\t%.8f\t%.8f\n" % (x, st.Rho[i], st.RhoU[i], st.E[i]))
    xa, rho_a, _, rhou_a, e_a = st.get_analytic_solution(final_time)
    ana = ["X\tRho\tRhoU\tE\n"]
    for i, x in enumerate(xa):
        ana.append("%.8f\t%.8f\t%.8f\t%.8f\n" % (x, rho_a[i], rhou_a[i], e_a[i]))
    return "".join(lines), "".join(ana)
