"""ORACLE -- test infrastructure, not product code.

CPU (numpy, float64) restatement of the reference's per-stage 2D Euler DFR right-hand side
and SSP-RK(5,4) step (`RungeKutta5SSP.StepWorker` and its callees).  It consumes the same
flat `Problem` the device library consumes and exists only to check the CUDA path.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.

Parity status: the Go toolchain is absent in the build container, so the reference cannot
be run to produce golden vectors.  This restatement is pinned by the reference's own
known-answer tests for the path (tests/test_oracle_kats.py), which fix
operators, edge table, freestream preservation, polynomial/vortex divergence, RT gradients
and the vertex merge.  Roe/Lax/Roe-ER values, Riemann BC values, the sensor, dissipation,
dt logic and after-N-steps fields are NOT pinned by any reference test: for those this file
follows the Go source statement by statement (same operation order, no re-association) and
the parity is "unpinned" beyond that -- see DESIGN.md.

The partitioning of the reference (goroutine shards) does not change any value: every element
and edge is processed by the same arithmetic whichever shard owns it, so the oracle works on
the global arrays.  Layout: field arrays are [rows = nodes, cols = elements].

Reference (all under model_problems/Euler2D/ unless noted):
  euler.go:420-653 StepWorker, :655-726 InitializeDT/RHSInternalPoints/SetRTFluxInternal,
  :864-918 GetSolutionGradientUsingRTElement, :945-1002 global/local dt, :1048-1086 merges;
  edges.go:151-491; fluxes.go:53-87,135-190,284-503; bcs.go:11-133; fluids.go:289-336;
  dissipation.go:219-346,397-412,491-542,606-622; DG2D/dfr_shock_capturing.go:142-166;
  isentropic_vortex/analytic_vortex.go:31-82.
"""
import math

import numpy as np

BC_None, BC_In, BC_Dirichlet, BC_Slip, BC_Far, BC_Wall, BC_Cyl, BC_Neuman, BC_Out, \
    BC_IVortex, BC_Periodic, BC_PeriodicReversed = range(12)
FLUX_Average, FLUX_LaxFriedrichs, FLUX_Roe, FLUX_RoeER = range(4)

# SSP54 coefficients, euler.go:511-563
RK_A = (0.391752226571890,)
RK1 = (0.444370493651235, 0.555629506348765, 0.368410593050371)
RK2 = (0.620101851488403, 0.379898148511597, 0.251891774271694)
RK3 = (0.178079954393132, 0.821920045606868, 0.544974750228521)
RK4 = (0.517231671970585, 0.096059710526146, 0.386708617503269, 0.063692468666290, 0.226007483236906)


def _pressure(gamma, rho, rho_u, rho_v, e):
    # fluids.go:289-320 (GetFlowFunctionBase, StaticPressure)
    oorho = 1.0 / rho
    u, v = rho_u * oorho, rho_v * oorho
    u2 = u * u + v * v
    q = 0.5 * rho * u2
    return (gamma - 1.0) * (e - q)


def _sound_speed(gamma, rho, rho_u, rho_v, e):
    oorho = 1.0 / rho
    p = _pressure(gamma, rho, rho_u, rho_v, e)
    return np.sqrt(np.abs(gamma * p * oorho))


def _velocity(rho, rho_u, rho_v):
    oorho = 1.0 / rho
    u, v = rho_u * oorho, rho_v * oorho
    return np.sqrt(u * u + v * v)


def flux_calc_base(gamma, rho, rho_u, rho_v, e):
    # fluxes.go:76-87
    oorho = 1.0 / rho
    u = rho_u * oorho
    v = rho_v * oorho
    p = _pressure(gamma, rho, rho_u, rho_v, e)
    fx = (rho_u, rho_u * u + p, rho_u * v, u * (e + p))
    fy = (rho_v, rho_v * u, rho_v * v + p, v * (e + p))
    return fx, fy


def ivortex_state(vortex, t, x, y):
    # isentropic_vortex/analytic_vortex.go:31-82
    beta, x0, y0, gamma, ufs = vortex
    oo2pi = 0.5 * (1.0 / math.pi)
    gm1 = gamma - 1.0
    oogm1 = 1.0 / gm1
    fac = 16.0 * gamma * (math.pi * math.pi)
    beta2 = beta * beta
    u, v = ufs, 0.0
    xmut, ymvt = x - u * t, y - v * t
    r2 = (xmut - x0) * (xmut - x0) + (ymvt - y0) * (ymvt - y0)
    ex1r = np.exp(1.0 - r2)
    tv1 = 1.0 - (gm1 * beta2 * np.exp(2.0 * (1.0 - r2)) / fac)
    u = u - beta * ex1r * (ymvt - y0) * oo2pi
    v = v + beta * ex1r * (xmut - x0) * oo2pi
    rho = np.power(tv1, oogm1)
    p = np.power(rho, gamma)
    q = 0.5 * rho * (u * u + v * v)
    oogm1c = 1.0 / (gamma - 1.0)
    return rho, rho * u, rho * v, p * oogm1c + q


def riemann_bc(fs, q_int, q_inf, nx, ny):
    """bcs.go:70-133.  fs = (Gamma, Qinf[4], Pinf, QQinf, Cinf, Alpha, Minf); arrays broadcast."""
    gamma, p_inf, c_inf, minf = fs[0], fs[5], fs[7], fs[9]
    r0, r1, r2, r3 = q_int
    rho_int, u_int, v_int = r0, r1 / r0, r2 / r0
    p_int = _pressure(gamma, r0, r1, r2, r3)
    c_int = _sound_speed(gamma, r0, r1, r2, r3)
    gm1 = gamma - 1.0
    oogm1 = 1.0 / gm1
    rho_inf, u_inf, v_inf = q_inf[0], q_inf[1] / q_inf[0], q_inf[2] / q_inf[0]
    tx, ty = -ny, nx
    vn_int = nx * u_int + ny * v_int
    inflow = vn_int < 0
    if minf <= 1.0:
        vn_inf = nx * u_inf + ny * v_inf
        rinf = vn_inf - 2.0 * c_inf * oogm1
        rint = vn_int + 2.0 * c_int * oogm1
        vnorm = 0.5 * (rint + rinf)
        c = 0.25 * gm1 * (rint - rinf)
        vtang_in = tx * u_inf + ty * v_inf
        beta_in = p_inf / np.power(rho_inf, gamma)
        vtang_out = tx * u_int + ty * v_int
        beta_out = p_int / np.power(rho_int, gamma)
        # NaN VnormInt matches neither case in the reference: Vtang = Beta = 0
        isnan = np.isnan(vn_int)
        vtang = np.where(inflow, vtang_in, np.where(isnan, 0.0, vtang_out))
        beta = np.where(inflow, beta_in, np.where(isnan, 0.0, beta_out))
        u = vnorm * nx + vtang * tx
        v = vnorm * ny + vtang * ty
        rho = np.power(c * c / (gamma * beta), oogm1)
        p = beta * np.power(rho, gamma)
        return [rho, rho * u, rho * v, p * oogm1 + 0.5 * rho * (u * u + v * v)]
    ones = np.ones_like(r0)
    return [np.where(inflow, q_inf[0] * ones, r0), np.where(inflow, q_inf[1] * ones, r1),
            np.where(inflow, q_inf[2] * ones, r2), np.where(inflow, q_inf[3] * ones, r3)]


def roe_flux(gamma, ql, qr, nx, ny):
    # fluxes.go:284-413
    gm1 = gamma - 1.0
    rho_ulr = ql[1] * nx + ql[2] * ny
    rho_vlr = ql[1] * (-ny) + ql[2] * nx
    rho_urr = qr[1] * nx + qr[2] * ny
    rho_vrr = qr[1] * (-ny) + qr[2] * nx
    rho_l, u_l, v_l = ql[0], rho_ulr / ql[0], rho_vlr / ql[0]
    rho_r, u_r, v_r = qr[0], rho_urr / qr[0], rho_vrr / qr[0]
    p_l = _pressure(gamma, ql[0], ql[1], ql[2], ql[3])
    p_r = _pressure(gamma, qr[0], qr[1], qr[2], qr[3])
    h_l, h_r = (ql[3] + p_l) / rho_l, (qr[3] + p_r) / rho_r
    rho_ls, rho_rs = np.sqrt(rho_l), np.sqrt(rho_r)
    rho_lsrs = rho_ls + rho_rs
    rho = rho_ls * rho_rs
    u = (rho_ls * u_l + rho_rs * u_r) / rho_lsrs
    v = (rho_ls * v_l + rho_rs * v_r) / rho_lsrs
    h = (rho_ls * h_l + rho_rs * h_r) / rho_lsrs
    c2 = gm1 * (h - 0.5 * (u * u + v * v))
    c = np.sqrt(c2)
    dw1 = -0.5 * (rho * (u_r - u_l)) / c + 0.5 * (p_r - p_l) / c2
    dw2 = (rho_r - rho_l) - (p_r - p_l) / c2
    dw3 = rho * (v_r - v_l)
    dw4 = 0.5 * (rho * (u_r - u_l)) / c + 0.5 * (p_r - p_l) / c2
    dw1 = np.abs(u - c) * dw1
    dw2 = np.abs(u) * dw2
    dw3 = np.abs(u) * dw3
    dw4 = np.abs(u + c) * dw4
    f0 = 0.5 * (rho_ulr + rho_urr)
    f1 = 0.5 * (rho_ulr * u_l + rho_urr * u_r + +p_l + p_r)
    f2 = 0.5 * (rho_vlr * u_l + rho_vrr * u_r)
    f3 = 0.5 * ((p_l + ql[3]) * u_l + (p_r + qr[3]) * u_r)
    f0 = f0 - 0.5 * (dw1 + dw2 + dw4)
    f1 = f1 - 0.5 * (dw1 * (u - c) + dw2 * u + dw4 * (u + c))
    f2 = f2 - 0.5 * (dw1 * v + dw2 * v + dw3 + dw4 * v)
    f3 = f3 - 0.5 * (dw1 * (h - u * c) + 0.5 * dw2 * (u * u + v * v) + dw3 * v + dw4 * (h + u * c))
    return [f0, nx * f1 - ny * f2, ny * f1 + nx * f2, f3]


def lax_flux(gamma, ql, qr, nx, ny):
    # fluxes.go:161-190
    rho_l, rho_r = ql[0], qr[0]
    u_l, v_l = ql[1] / rho_l, ql[2] / rho_l
    u_r, v_r = qr[1] / rho_r, qr[2] / rho_r
    p_l, p_r = _pressure(gamma, *ql), _pressure(gamma, *qr)
    c_l, c_r = _sound_speed(gamma, *ql), _sound_speed(gamma, *qr)
    max_v = np.maximum(np.sqrt(u_l * u_l + v_l * v_l) + c_l, np.sqrt(u_r * u_r + v_r * v_r) + c_r)
    f = [0.5 * (nx * (ql[1] + qr[1]) + ny * (ql[2] + qr[2])),
         0.5 * (nx * (ql[1] * u_l + qr[1] * u_r + p_l + p_r) + ny * (ql[1] * v_l + qr[1] * v_r)),
         0.5 * (nx * (ql[2] * u_l + qr[2] * u_r) + ny * (ql[2] * v_l + qr[2] * v_r + p_l + p_r)),
         0.5 * (nx * ((p_l + ql[3]) * u_l + (p_r + qr[3]) * u_r) + ny * ((p_l + ql[3]) * v_l + (p_r + qr[3]) * v_r))]
    return [f[n] + 0.5 * max_v * (ql[n] - qr[n]) for n in range(4)]


def avg_flux(gamma, ql, qr, nx, ny, flux_calc=None):
    # fluxes.go:135-159 (goes through the mockable CalculateFlux, fluxes.go:71-74)
    if flux_calc is None:
        flux_calc = lambda *q: flux_calc_base(gamma, *q)      # noqa: E731
    fxl, fyl = flux_calc(*ql)
    fxr, fyr = flux_calc(*qr)
    return [nx * (0.5 * (fxl[n] + fxr[n])) + ny * (0.5 * (fyl[n] + fyr[n])) for n in range(4)]


def roe_er_flux(gamma, ql, qr, nx, ny):
    # fluxes.go:415-503 -- reproduced as written, including `(dPu+dPp)*ny` in the energy row
    # and h = (HL+HR)*ooRs with HL = EL+pL (not divided by rho).
    gm1 = gamma - 1.0
    rho_l, rho_r = ql[0], qr[0]
    oorho_l, oorho_r = 1.0 / rho_l, 1.0 / rho_r
    rho_ls, rho_rs = np.sqrt(rho_l), np.sqrt(rho_r)
    u_l, v_l = ql[1] * oorho_l, ql[2] * oorho_l
    u_r, v_r = qr[1] * oorho_r, qr[2] * oorho_r
    e_l, e_r = ql[3], qr[3]
    uu_l, uu_r = nx * u_l + ny * v_l, nx * u_r + ny * v_r
    p_l, p_r = _pressure(gamma, *ql), _pressure(gamma, *qr)
    hh_l, hh_r = e_l + p_l, e_r + p_r
    oors = 1.0 / (rho_ls + rho_rs)
    u, v, h = (rho_ls * u_l + rho_rs * u_r) * oors, (rho_ls * v_l + rho_rs * v_r) * oors, (hh_l + hh_r) * oors
    rho = rho_ls * rho_rs
    hh = h * rho
    uu = nx * u + ny * v
    c2 = gm1 * (h - 0.5 * (u * u + v * v))
    c = np.sqrt(c2)
    ooc = 1.0 / c
    uabs = np.abs(uu)
    f = [0.5 * (uu_l * rho_l + uu_r * rho_r),
         0.5 * (uu_l * rho_l * u_l + p_l * nx + uu_r * rho_r * u_r + p_r * nx),
         0.5 * (uu_l * rho_l * v_l + p_l * ny + uu_r * rho_r * v_r + p_r * ny),
         0.5 * (uu_l * hh_l + uu_r * hh_r)]
    uef = 0.05 * c
    du, dv = (u_r - u_l), (v_r - v_l)
    delta_v2 = du * du + dv * dv
    with np.errstate(divide="ignore", invalid="ignore"):     # u = v = 0 (tube at rest): the `small` branch is taken
        oovmag = 1.0 / np.sqrt(u * u + v * v)
        small = delta_v2 < 0.01 * c2
        n1x = np.where(small, nx, oovmag * du)
        n1y = np.where(small, ny, oovmag * dv)
    n2x, n2y = n1y * (nx * n1y - n1x * ny), -n1x * (nx * n1y - n1x * ny)
    alp1, alp2 = nx * n1x + ny * n1y, nx * n2x + ny * n2y
    u1x, u1y = n1x * u, n1y * v
    u2x, u2y = n2x * u, n2y * v
    urot = np.sqrt(alp1 * alp1 * (u1x * u1x + u1y * u1y)) + np.sqrt(alp2 * alp2 * (u2x * u2x + u2y * u2y))
    sigma = np.maximum(uabs, np.minimum(uef, urot))
    sgn = lambda a: np.copysign(a, 1.0)      # noqa: E731  math.Copysign(x, 1) == |x|
    uabs_prime = uabs - 0.25 * np.maximum(0.0, uu_r - uu_l) * (sgn(uu + c) - sgn(uu - c))
    d_u, d_p, d_rho = uu_r - uu_l, p_r - p_l, rho_r - rho_l
    d_rho_u, d_rho_v, d_e = rho_r * u_r - rho_l * u_l, rho_r * v_r - rho_l * v_l, e_r - e_l
    d_pu = rho * d_u * np.maximum(0.0, c - uabs_prime)
    swt = sgn(uu) * np.minimum(uabs_prime, c)
    d_pp = swt * d_p * ooc
    d_uu = swt * d_u * ooc
    f[0] = f[0] - 0.5 * (sigma * d_rho + (d_pu + d_pp) * 0 + d_uu * rho)
    f[1] = f[1] - 0.5 * (sigma * d_rho_u + (d_pu + d_pp) * nx + d_uu * rho * u)
    f[2] = f[2] - 0.5 * (sigma * d_rho_v + (d_pu + d_pp) * ny + d_uu * rho * v)
    f[3] = f[3] - 0.5 * (sigma * d_e + (d_pu + d_pp) * ny + d_uu * hh)
    return f


# FlowFunction numbering, fluids.go:209-223
FF_Density, FF_XMomentum, FF_YMomentum, FF_Energy, FF_Mach, FF_StaticPressure, FF_DynamicPressure, \
    FF_PressureCoefficient, FF_SoundSpeed, FF_Velocity, FF_XVelocity, FF_YVelocity, FF_Enthalpy, FF_Entropy = range(14)


def flow_function(fs, q, pf):
    """FreeStream.GetFlowFunctionBase (fluids.go:289-336) on arrays; fs = FreeStream.as_array() tuple."""
    gamma, p_inf, qq_inf = fs[0], fs[5], fs[6]
    rho, rho_u, rho_v, e = q
    gm1 = gamma - 1.0
    oorho = 1.0 / rho
    if pf <= FF_Energy:
        return np.array(q[pf], dtype=np.float64, copy=True)
    if pf == FF_XVelocity:
        return rho_u * oorho
    if pf == FF_YVelocity:
        return rho_v * oorho
    u, v = rho_u * oorho, rho_v * oorho
    u2 = u * u + v * v
    qq = 0.5 * rho * u2
    p = gm1 * (e - qq)
    if pf == FF_Velocity:
        return np.sqrt(u2)
    if pf == FF_DynamicPressure:
        return qq
    if pf == FF_StaticPressure:
        return p
    if pf == FF_PressureCoefficient:
        return (p - p_inf) / qq_inf
    if pf == FF_SoundSpeed:
        return np.sqrt(np.abs(gamma * p * oorho))
    if pf == FF_Enthalpy:
        return (e + p) / rho
    if pf == FF_Entropy:
        return np.log(p) - gamma * np.log(rho)
    if pf == FF_Mach:
        return np.sqrt(u2) / np.sqrt(np.abs(gamma * p * oorho))
    raise ValueError("flow function %d does not go through GetFlowFunction" % pf)


FF_ShockFunction, FF_EpsilonDissipation, FF_EpsilonDissipationC0 = 100, 101, 102       # fluids.go:224-226


def shock_indicator(p, rho, kappa):
    """ModeAliasShockFinder.ShockIndicator per element (DG2D/dfr_shock_capturing.go:189-235): moment of U - Clipper U with
    the mass-matrix diagonal, Persson ramp with S0 = 4 / N^4.  rho = [NpInt, K]."""
    clipper = np.eye(p.NpInt) - p.D
    t1 = rho - clipper @ rho
    mass = np.diag(p.MassMatrix)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        se = np.log10((mass * t1 * t1).sum(axis=0) / (mass * rho * rho).sum(axis=0))
        s0 = 4.0 / float(p.N) ** 4 if p.N > 0 else np.inf
        left, right = s0 - kappa, s0 + kappa
        sigma = np.zeros(rho.shape[1])
        mid = (se >= left) & (se <= right)
        sigma[mid] = 0.5 * (1.0 + np.sin(np.pi * (0.5 / kappa) * (se[mid] - s0)))
        sigma[se > right] = 1.0
    return sigma


def plot_field(p, q, pf, graph_interp):
    """Euler.GetPlotField for the GetFlowFunction family (plot.go:14-86): node values, GraphInterp product,
    AverageGraphFieldVertices (DG2D/graphics_support2.go:184-199), transpose -> [K, NpGraph]; the AVS writer then
    narrows to float32 (DG2D/graphics_support.go:80-93), which the caller applies."""
    if pf == FF_ShockFunction:
        # plot.go:30-47: c.ShockFinder (Kappa = ip.Kappa as given, euler.go:82) with the limiter, else NewAliasShockFinder(2)
        kappa = p.Kappa if p.Dissipation else 2.0
        fld = np.tile(shock_indicator(p, q[0], kappa), (p.NpInt, 1))
    else:
        fld = flow_function(tuple(p.FSFar.as_array()), q, pf)
    field = graph_interp @ fld
    npe = p.NpEdge + 2
    for n_edge in range(3):
        iv = n_edge * (npe - 1)
        ivp = iv + 1
        ivm = 3 * (npe - 1) - 1 if n_edge == 0 else n_edge * (npe - 1) - 1
        field[iv] = 0.5 * (field[ivp] + field[ivm])
    return np.ascontiguousarray(field.T)


_FLUX_FUNCS = {FLUX_Average: avg_flux, FLUX_LaxFriedrichs: lax_flux, FLUX_Roe: roe_flux, FLUX_RoeER: roe_er_flux}


class OracleSolver:
    """Same surface as gocfd_b200.lib.Dfr2d: set_state / step / rhs / residual / get_state."""

    def __init__(self, p, flux_calc=None):
        self.p = p
        self.N, self.K, self.NE = p.N, p.K, p.NE
        self.NpInt, self.NpEdge, self.NpFlux = p.NpInt, p.NpEdge, p.NpFlux
        self.gamma = p.Gamma
        # freestream tuples indexed like as_array(): [0]=gamma, [5]=Pinf, [7]=Cinf, [9]=Minf
        self.fs = {}
        for name, fs in (("far", p.FSFar), ("in", p.FSIn), ("out", p.FSOut)):
            a = fs.as_array()
            self.fs[name] = (a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9])
        self.vortex = tuple(p.Vortex.as_array())
        self.flux_calc = flux_calc or (lambda r, ru, rv, e: flux_calc_base(self.gamma, r, ru, rv, e))
        k, ni = self.K, self.NpInt
        self.Q = [np.zeros((4, ni, k)) for _ in range(5)]          # c.Q, Q1..Q4
        self.Residual = np.zeros((4, ni, k))
        self.RHSQ = np.zeros((4, ni, k))
        self.Q_Face = np.zeros((4, 3 * self.NpEdge, k))
        self.F_RT_DOF = np.zeros((4, self.NpFlux, k))
        self.DT = np.zeros(k)
        self.DTVisc = np.zeros(k)
        self.EdgeFlux = np.zeros((3, 4, self.NE, self.NpEdge))      # [type][n][edge][i]
        self.Aggregates = np.zeros((self.NE, 2))
        self.Time, self.GlobalDT, self.StepCount = 0.0, 0.0, 0
        self.hK = p.EdgeLenMax / float((p.N + 1) * (p.N + 1))
        # row index helpers per edge
        i = np.arange(self.NpEdge)
        self.rowsL = p.edge_numL[:, None] * self.NpEdge + i[None, :]
        self.rowsR = p.edge_numR[:, None] * self.NpEdge + (self.NpEdge - 1 - i)[None, :]
        self.kLc = p.edge_kL[:, None].astype(np.int64)
        self.kRc = np.maximum(p.edge_kR, 0)[:, None].astype(np.int64)
        self.nxL = p.FaceNormX[p.edge_numL, p.edge_kL]
        self.nyL = p.FaceNormY[p.edge_numL, p.edge_kL]
        self.bp_of_edge = np.full(self.NE, -1, dtype=np.int64)
        self.bp_of_edge[p.bp_edge] = np.arange(p.NBP)
        # DXMetric / DYMetric (CalculateRTBasedDerivativeMetrics, DG2D/dfr_startup.go:213-254): built by NewDFR2D for every
        # run -- the PerssonC0 limiter and the gradient plot fields (plot.go:54-77) both use them
        ne = self.NpEdge
        self.DXMetric = np.empty((self.NpFlux, k))
        self.DYMetric = np.empty((self.NpFlux, k))
        self.DXMetric[:ni], self.DXMetric[ni:2 * ni] = p.Jinv[:, 0], p.Jinv[:, 2]
        self.DYMetric[:ni], self.DYMetric[ni:2 * ni] = p.Jinv[:, 1], p.Jinv[:, 3]
        oojd = 1.0 / p.Jdet
        for fn in range(3):
            rows = slice(2 * ni + fn * ne, 2 * ni + (fn + 1) * ne)
            self.DXMetric[rows] = oojd * p.FaceNormX[fn] * p.IInII[fn]
            self.DYMetric[rows] = oojd * p.FaceNormY[fn] * p.IInII[fn]
        if p.Dissipation:
            self.sd_kappa = p.Kappa if p.Kappa != 0.0 else 5.0      # dissipation.go:140-147
            self.Eps0 = 5.0 / 1.5
            nv = p.NV
            self.SigmaScalar = np.zeros(k)
            self.EpsilonScalar = np.zeros(k)
            self.SigmaVertex = np.zeros(nv)
            self.EpsVertex = np.zeros(nv)
            self.Epsilon = np.zeros((self.NpFlux, k))
            self.Se = np.zeros(k)
            self.DissX = np.zeros((4, self.NpFlux, k))
            self.DissY = np.zeros((4, self.NpFlux, k))

    # ---- state I/O -----------------------------------------------------------------
    def set_state(self, q):
        self.Q[0][...] = q

    def get_state(self):
        return self.Q[0].copy()

    def residual(self):
        return [float(self.Residual[n].max()) for n in range(4)]

    # ---- phases --------------------------------------------------------------------
    def interpolate_to_edges(self, q):
        for n in range(4):
            self.Q_Face[n] = self.p.FluxEdgeInterp @ q[n]

    def update_se_moment(self, rho):
        p = self.p
        x, y, dq = p.P @ rho, p.MassMatrix @ rho, p.D @ rho
        num = np.zeros(self.K)
        den = np.zeros(self.K)
        for i in range(self.NpInt):
            num += dq[i] * x[i]
            den += rho[i] * y[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            self.Se = np.log10(num / den)

    def update_shock_finder_sigma(self):
        kappa = self.sd_kappa
        s0 = 4.0 / math.pow(float(self.N + 1), 4.0)
        left, right = s0 - kappa, s0 + kappa
        ookappa = 0.5 / kappa
        se = self.Se
        with np.errstate(invalid="ignore"):
            mid = 0.5 * (1.0 + np.sin(math.pi * ookappa * (se - s0)))
            sigma = np.where(se < left, 0.0, np.where(se <= right, mid, 1.0))
        nan = np.isnan(se)
        if nan.any():   # `sigma` is declared outside the loop: a NaN Se inherits the previous element's value
            for k in np.flatnonzero(nan):
                sigma[k] = sigma[k - 1] if k > 0 else 0.0
        self.SigmaScalar = sigma

    def calculate_element_viscosity(self):
        self.EpsilonScalar = self.Eps0 * self.hK * self.SigmaScalar

    def merge_to_vertices(self):
        ev = self.p.EToV.reshape(-1)
        for src, dst in ((self.SigmaScalar, self.SigmaVertex), (self.EpsilonScalar, self.EpsVertex)):
            vals = np.repeat(src, 3)
            touched = np.zeros(dst.shape[0], dtype=bool)
            touched[ev] = True
            acc = np.full(dst.shape[0], -np.inf)
            np.maximum.at(acc, ev, vals)
            dst[touched] = acc[touched]

    def merge_vertex_sigma_to_element(self):
        ev = self.p.EToV
        acc = 0.0 + self.SigmaVertex[ev[:, 0]]
        acc = acc + self.SigmaVertex[ev[:, 1]]
        acc = acc + self.SigmaVertex[ev[:, 2]]
        self.SigmaScalar = acc / 3.0

    def interpolate_epsilon(self):
        ev = self.p.EToV
        vals = np.stack([self.EpsVertex[ev[:, 0]], self.EpsVertex[ev[:, 1]], self.EpsVertex[ev[:, 2]]])
        self.Epsilon = self.p.Bary @ vals

    def limit_filter_solution(self, q):
        p = self.p
        alpha = np.sin(0.5 * math.pi * self.SigmaScalar)
        for n in range(4):
            uh = p.Vinv @ q[n]
            for i in range(1, self.NpInt):
                uh[i] *= p.ModeFilter[i] * (1.0 - alpha)
            q[n] = p.V @ uh

    def _gather_lr(self, sel):
        ql = [self.Q_Face[n][self.rowsL[sel], self.kLc[sel]] for n in range(4)]
        qr = [self.Q_Face[n][self.rowsR[sel], self.kRc[sel]] for n in range(4)]
        return ql, qr

    def calculate_edge_euler_flux(self, t):
        p = self.p
        eq, nf = self.EdgeFlux[1], self.EdgeFlux[0]
        all_e = np.arange(self.NE)
        for n in range(4):      # EdgeQValues: owner side, before any BC overwrite (edges.go:344-350)
            eq[n] = self.Q_Face[n][self.rowsL, self.kLc]
        nf[...] = 0.0
        shared = p.edge_nconn == 2
        if shared.any():
            ql, qr = self._gather_lr(shared)
            nxs, nys = self.nxL[shared][:, None], self.nyL[shared][:, None]
            if p.FluxType == FLUX_Average:
                f = avg_flux(self.gamma, ql, qr, nxs, nys, self.flux_calc)
            else:
                f = _FLUX_FUNCS[p.FluxType](self.gamma, ql, qr, nxs, nys)
            for n in range(4):
                nf[n][shared] = f[n]
        bnd = all_e[p.edge_nconn == 1]
        for bc in np.unique(p.edge_bc[bnd]):
            sel = bnd[p.edge_bc[bnd] == bc]
            nx, ny = self.nxL[sel][:, None], self.nyL[sel][:, None]
            rows, kk = self.rowsL[sel], self.kLc[sel]
            q_int = [self.Q_Face[n][rows, kk] for n in range(4)]
            if bc in (BC_Periodic, BC_PeriodicReversed):
                continue
            if bc in (BC_Wall, BC_Cyl):
                pw = _pressure(self.gamma, *q_int)
                nf[1][sel], nf[2][sel] = nx * pw, ny * pw
                continue
            if bc in (BC_Far, BC_In, BC_Out):
                fs = self.fs[{BC_Far: "far", BC_In: "in", BC_Out: "out"}[bc]]
                q_int = riemann_bc(fs, q_int, fs[1:5], nx, ny)
            elif bc == BC_IVortex:
                b = self.bp_of_edge[sel]
                q_ex = ivortex_state(self.vortex, t, p.bp_x[b], p.bp_y[b])
                q_int = riemann_bc(self.fs["far"], q_int, q_ex, nx, ny)
            if bc != BC_None:
                for n in range(4):      # BCs overwrite Q_Face in place (bcs.go:47-50, 63-66)
                    self.Q_Face[n][rows, kk] = q_int[n]
            fx, fy = self.flux_calc(*q_int)
            for n in range(4):
                nf[n][sel] = nx * fx[n] + ny * fy[n]

    def store_edge_aggregates(self):
        p = self.p
        oohk = (1.0 / self.hK)[p.edge_kL][:, None]
        q = [self.Q_Face[n][self.rowsL, self.kLc] for n in range(4)]
        c = _sound_speed(self.gamma, *q)
        u = _velocity(q[0], q[1], q[2])
        self.Aggregates[:, 0] = (oohk * (u + c)).max(axis=1)
        if p.Dissipation:
            eps = self.Epsilon[2 * self.NpInt + self.rowsL, self.kLc]
            self.Aggregates[:, 1] = (oohk * oohk * eps).max(axis=1)

    def calculate_epsilon_gradient(self, q):
        for n in range(4):
            gx, gy = self.solution_gradient_rt(n, q)
            self.DissX[n] = gx * self.Epsilon
            self.DissY[n] = gy * self.Epsilon

    def solution_gradient_rt(self, n, q):
        """GetSolutionGradientUsingRTElement (euler.go:864-918) for conserved variable n: interior rows from q, edge rows
        from the EdgeQValues store (owner side of the last CalculateEdgeEulerFlux, reversed for the neighbour), times
        DXMetric / DYMetric, then the RT divergence.  Returns GradX, GradY [NpFlux, K]."""
        p = self.p
        ni, ne = self.NpInt, self.NpEdge
        eq = self.EdgeFlux[1]
        un = np.empty((self.NpFlux, self.K))
        un[:ni] = q[n]
        un[ni:2 * ni] = q[n]
        for e in range(3):
            ei = p.EtoEdge[:, e]
            owner = p.edge_kL[ei] == np.arange(self.K)
            vals = eq[n][ei]                                   # [K, NpEdge] in owner order
            vals = np.where(owner[:, None], vals, vals[:, ::-1])
            un[2 * ni + e * ne:2 * ni + (e + 1) * ne] = vals.T
        return p.Div @ (self.DXMetric * un), p.Div @ (self.DYMetric * un)

    def gradient_plot_field(self, pf):
        """GetPlotField for XGradientDensity..YGradientEnergy (plot.go:54-77; fluids.go:227-234: 200+n = R direction,
        300+n = S direction): the RT gradient of variable n of the CURRENT c.Q with the edge values the store holds from
        the last stage that ran (the input of stage 5 of the last step; zeros before the first step) -- not interpolated,
        [NpFlux, K] doubles."""
        if not (200 <= pf <= 203 or 300 <= pf <= 303):
            raise ValueError("not a gradient plot field: %d" % pf)
        gx, gy = self.solution_gradient_rt(pf % 100, self.Q[0])
        return gx if pf < 300 else gy

    def store_edge_viscous_flux(self):
        p = self.p
        ni = self.NpInt
        vf = self.EdgeFlux[2]
        eq = self.EdgeFlux[1]
        rl = 2 * ni + self.rowsL
        rr = 2 * ni + self.rowsR
        nx, ny = self.nxL[:, None], self.nyL[:, None]
        shared = (p.edge_nconn == 2)[:, None]
        omega = 1.0 * float(self.N * self.N)
        ooel = (1.0 / p.edge_len)[:, None]
        lam = 0.5 * (self.Epsilon[rl, self.kLc] + self.Epsilon[rr, self.kRc])
        for n in range(4):
            vfl = nx * self.DissX[n][rl, self.kLc] + ny * self.DissY[n][rl, self.kLc]
            vfr = nx * self.DissX[n][rr, self.kRc] + ny * self.DissY[n][rr, self.kRc]   # normalR := normalL
            inner = 0.5 * (vfl + vfr)
            # the "jump" uses the same stored owner-side slice on both sides (edges.go:225-236)
            inner = inner - (omega * lam * ooel) * (eq[n] - eq[n][:, ::-1])
            vf[n] = np.where(shared, inner, vfl)

    def calc_element_max_wave_speed(self, rk):
        p = self.p
        if rk == 0:
            self.DT[:] = -100.0
        gmax, gmaxv = -np.finfo(np.float64).max, -np.finfo(np.float64).max
        for e in range(3):
            agg = self.Aggregates[p.EtoEdge[:, e]]
            self.DT = np.maximum(self.DT, agg[:, 0])
            gmax = max(gmax, float(agg[:, 0].max()))
            if p.Dissipation:
                self.DTVisc = np.maximum(self.DTVisc, agg[:, 1])
                gmaxv = max(gmaxv, float(agg[:, 1].max()))
        return gmax, gmaxv

    def calculate_global_dt(self, gmax, gmaxv):
        p = self.p
        with np.errstate(divide="ignore"):
            self.GlobalDT = float(np.float64(p.CFL) / np.float64(max(0.0, gmax)))
        if p.Dissipation:
            c_diff = 1.0 / float((self.N + 1) * (self.N + 1))
            gv = max(0.0, gmaxv)
            self.GlobalDT = min(self.GlobalDT, c_diff / gv if gv != 0.0 else math.inf)
        if self.Time + self.GlobalDT > p.FinalTime:
            self.GlobalDT = p.FinalTime - self.Time

    def calculate_local_dt(self):
        p = self.p
        if not p.LocalTimeStepping:
            self.DT[:] = self.GlobalDT
            return
        self.DT = p.CFL / self.DT
        if p.Dissipation:
            c_diff = 1.0 / float((self.N + 1) * (self.N + 1))
            m = self.DTVisc > 1.0e-9
            self.DTVisc[m] = c_diff / self.DTVisc[m]
            self.DT[m] = np.minimum(self.DT[m], self.DTVisc[m])

    def set_rt_flux_internal(self, q):
        p = self.p
        ni = self.NpInt
        self.F_RT_DOF[...] = 0.0
        fx, fy = self.flux_calc(q[0], q[1], q[2], q[3])
        jd = p.Jdet[None, :]
        j0, j1, j2, j3 = (p.Jinv[:, c][None, :] for c in range(4))
        for n in range(4):
            self.F_RT_DOF[n][:ni] = jd * (j0 * fx[n] + j1 * fy[n])
            self.F_RT_DOF[n][ni:2 * ni] = jd * (j2 * fx[n] + j3 * fy[n])

    def _edge_rows_from_store(self, store, signed):
        """[3*NpEdge, K] rows for every element from an edge store: owner order, or reversed
        (and negated when `signed`) for the neighbour (edges.go:93-113, :469-479)."""
        p = self.p
        ne = self.NpEdge
        out = np.empty((3 * ne, self.K))
        for e in range(3):
            ei = p.EtoEdge[:, e]
            owner = (p.edge_kL[ei] == np.arange(self.K))[:, None]
            vals = store[ei]
            rev = vals[:, ::-1]
            out[e * ne:(e + 1) * ne] = np.where(owner, vals, -rev if signed else rev).T
        return out

    def set_rt_flux_on_edges(self):
        p = self.p
        ni, ne = self.NpInt, self.NpEdge
        iin = np.repeat(p.IInII, ne, axis=0)
        for n in range(4):
            self.F_RT_DOF[n][2 * ni:] = self._edge_rows_from_store(self.EdgeFlux[0][n], True) * iin

    def rhs_internal_points(self):
        p = self.p
        oojd = (1.0 / p.Jdet)[None, :]
        for n in range(4):
            self.RHSQ[n] = (p.DivInt @ self.F_RT_DOF[n]) * (-oojd)

    def add_dissipation(self):
        p = self.p
        ni, ne = self.NpInt, self.NpEdge
        jd = p.Jdet[None, :]
        j0, j1, j2, j3 = (p.Jinv[:, c][None, :] for c in range(4))
        iin = np.repeat(p.IInII, ne, axis=0)
        ei = [p.EtoEdge[:, e] for e in range(3)]
        sign = np.repeat(np.stack([np.where(p.edge_kL[ei[e]] == np.arange(self.K), 1.0, -1.0)
                                   for e in range(3)]), ne, axis=0)
        oojd = (1.0 / p.Jdet)[None, :]
        for n in range(4):
            dof = np.empty((self.NpFlux, self.K))
            dx, dy = self.DissX[n][:ni], self.DissY[n][:ni]
            dof[:ni] = jd * (j0 * dx + j1 * dy)
            dof[ni:2 * ni] = jd * (j2 * dx + j3 * dy)
            dof[2 * ni:] = self._edge_rows_from_store(self.EdgeFlux[2][n], False) * iin * sign
            self.RHSQ[n] += oojd * (p.DivInt @ dof)

    def rk_advance(self, rk):
        p = self.p
        q0, q1, q2, q3, q4 = self.Q
        qqq = self.Q[rk]
        self.set_rt_flux_internal(qqq)
        self.set_rt_flux_on_edges()
        self.rhs_internal_points()
        if p.Dissipation:
            self.add_dissipation()
            self.limit_filter_solution(self.RHSQ)
        dt = self.DT[None, :]
        for n in range(4):
            rhs = self.RHSQ[n]
            dt_rhs = dt * rhs
            if rk == 0:
                q1[n] = q0[n] + RK_A[0] * dt_rhs
            elif rk == 1:
                q2[n] = RK1[0] * q0[n] + RK1[1] * q1[n] + RK1[2] * dt_rhs
            elif rk == 2:
                q3[n] = RK2[0] * q0[n] + RK2[1] * q2[n] + RK2[2] * dt_rhs
            elif rk == 3:
                self.Residual[n] = rhs.copy()
                q4[n] = RK3[0] * q0[n] + RK3[1] * q3[n] + RK3[2] * dt_rhs
            else:
                dt_r3 = dt * self.Residual[n]
                r = (-q0[n] + RK4[0] * q2[n] + RK4[1] * q3[n] + RK4[2] * q4[n]
                     + RK4[3] * dt_r3 + RK4[4] * dt_rhs)
                self.Residual[n] = r
                q0[n] = q0[n] + r

    # ---- one RK stage: StepWorker (euler.go:420-653) --------------------------------
    def stage(self, rk):
        p = self.p
        qqq = self.Q[rk]
        if p.Dissipation:
            self.update_se_moment(qqq[0])
            self.update_shock_finder_sigma()
            self.calculate_element_viscosity()
            self.merge_to_vertices()
            self.merge_vertex_sigma_to_element()
            self.interpolate_epsilon()
            if rk == 2:
                self.limit_filter_solution(qqq)
        self.interpolate_to_edges(qqq)
        self.calculate_edge_euler_flux(self.Time)
        if p.Dissipation:
            self.calculate_epsilon_gradient(qqq)
        self.store_edge_aggregates()
        if p.Dissipation:
            self.store_edge_viscous_flux()
        gmax, gmaxv = self.calc_element_max_wave_speed(rk)
        if not p.LocalTimeStepping:
            self.calculate_global_dt(gmax, gmaxv)
        self.calculate_local_dt()
        self.rk_advance(rk)

    def rhs(self, rk=0):
        """RHSQ of stage `rk` evaluated on the current register of that stage (test hook)."""
        saved = ([q.copy() for q in self.Q], self.Residual.copy(), self.DT.copy(), self.DTVisc.copy(),
                 self.GlobalDT)
        self.stage(rk)
        out = self.RHSQ.copy()
        self.Q, self.Residual, self.DT, self.DTVisc, self.GlobalDT = saved
        return out

    def step(self, nsteps=1):
        p = self.p
        finished = False
        for _ in range(nsteps):
            for rk in range(5):
                self.stage(rk)
            self.Time += self.GlobalDT
            self.StepCount += 1
            finished = self.Time >= p.FinalTime or self.StepCount >= p.MaxIterations
            if finished:
                break
        return {"time": self.Time, "dt": self.GlobalDT, "steps": self.StepCount, "finished": finished}
