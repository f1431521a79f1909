"""`Euler.OutputFinal` (euler.go:221-318) on the host: what the reference writes once after the
last step, from the state read back with dfr2d_get_state.

  * SHOCKTUBE: shocktube.dat + shocktube_analytic.dat (sod_shock_tube.py).
  * FREESTREAM with PlotFields: plotfile.dat -- for every wall edge the RT edge-point coordinates and the requested
    flow functions (e.g. "Pressure Coefficient") at those points, "%.5e" columns.  The reference takes the values
    from `GetPlotField` (plot.go:14-86) at graph node `edgeNum*(NpEdge+1) + 1 + i`; here the same product
    GraphInterp . f(Q) is evaluated in float64 for the wall elements only (a few hundred columns) instead of the whole
    field.  The reference walks a Go map of edges, so its row order is unspecified; rows here follow the edge table.
  * `BestMatchFlowFunction` (fluids.go:128-160): token scoring of a field name against the flow-function names.  Go
    iterates a map, so a tie is broken at random there; here the lowest flow-function number wins.
"""
import os

import numpy as np

from gocfd_b200.host.readfiles import BC_Wall
from .sod_shock_tube import SODShockTube, shocktube_files

# FlowFunction numbering of fluids.go:199-232 (the GetFlowFunction family is 0..13)
FLOW_FUNCTIONS = [
    (0, "Density"), (1, "XMomentum"), (2, "YMomentum"), (3, "Energy"), (4, "Mach"), (5, "Static Pressure"),
    (6, "Dynamic Pressure"), (7, "Pressure Coefficient"), (8, "Sound Speed"), (9, "Velocity"), (10, "XVelocity"),
    (11, "YVelocity"), (12, "Enthalpy"), (13, "Entropy"),
    (100, "ShockFunction"), (101, "Artificial Dissipation Epsilon"), (102, "Artificial Dissipation Epsilon C0"),
    (200, "R Direction Gradient of Density"), (201, "R Direction Gradient of R Momentum"),
    (202, "R Direction Gradient of S Momentum"), (203, "R Direction Gradient of Energy"),
    (300, "S Direction Gradient of Density"), (301, "S Direction Gradient of R Momentum"),
    (302, "S Direction Gradient of S Momentum"), (303, "S Direction Gradient of Energy"),
]


def tokenize(s):
    return s.lower().replace("-", " ").replace("_", " ").split()


def best_match_flow_function(name):
    """(flow function number, matched) -- BestMatchFlowFunction (fluids.go:137-160)."""
    tokens_in = tokenize(name)
    best, best_score = -1, -1
    for ff, label in FLOW_FUNCTIONS:
        tokens = tokenize(label)
        score = 0
        for t_in in tokens_in:
            if any(t_in in t for t in tokens):
                score += 1
        if score > best_score:
            best, best_score = ff, score
    return best, best_score > 0


def get_flow_function(fs, q, pf):
    """FreeStream.GetFlowFunctionBase (fluids.go:289-336) on arrays q = (rho, rhoU, rhoV, E)."""
    rho, rho_u, rho_v, e = q
    if pf <= 3:
        return np.array(q[pf], dtype=np.float64, copy=True)
    gm1 = fs.Gamma - 1.0
    oorho = 1.0 / rho
    if pf == 10:
        return rho_u * oorho
    if pf == 11:
        return rho_v * oorho
    u, v = rho_u * oorho, rho_v * oorho
    u2 = u * u + v * v
    qq = 0.5 * rho * u2
    p = gm1 * (e - qq)
    if pf == 9:
        return np.sqrt(u2)
    if pf == 6:
        return qq
    if pf == 5:
        return p
    if pf == 7:
        return (p - fs.Pinf) / fs.QQinf
    if pf == 8:
        return np.sqrt(np.abs(fs.Gamma * p * oorho))
    if pf == 12:
        return (e + p) / rho
    if pf == 13:
        return np.log(p) - fs.Gamma * np.log(rho)
    if pf == 4:
        return np.sqrt(u2) / np.sqrt(np.abs(fs.Gamma * p * oorho))
    raise ValueError("flow function %d is not in the GetFlowFunction family" % pf)


def plot_field_elements(c, q, pf, elems):
    """GetPlotField (plot.go:14-86) restricted to the element columns `elems`: [len(elems), NpGraph] float64."""
    dfr = c.DFR
    fld = get_flow_function(c.FSFar, [q[n][:, elems] for n in range(4)], pf)
    field = dfr.graph_interp() @ fld
    npe = dfr.FluxElement.NpEdge + 2
    for n_edge in range(3):                      # AverageGraphFieldVertices (DG2D/graphics_support2.go:184-199)
        iv = n_edge * (npe - 1)
        ivm = 3 * (npe - 1) - 1 if n_edge == 0 else n_edge * (npe - 1) - 1
        field[iv] = 0.5 * (field[iv + 1] + field[ivm])
    return field.T


def wall_plot_data(c, q, field_names=None):
    """Text of plotfile.dat (euler.go:252-316) and (wall edges, points per edge)."""
    names = list(c.ip.PlotFields if field_names is None else field_names)
    dfr = c.DFR
    t = dfr.Tris
    rt = dfr.FluxElement
    ni, ned = rt.NpInt, rt.NpEdge
    wall = np.nonzero(t.bcType == BC_Wall)[0]
    k_wall = t.kL[wall]
    fields = []
    for name in names:
        ff, match = best_match_flow_function(name)
        if not match:
            raise RuntimeError("Unable to find matching flow function named: " + name)
        if ff > 13:
            raise NotImplementedError("wall output of '%s' is not a GetFlowFunction field" % name)
        fields.append(plot_field_elements(c, q, ff, k_wall))
    fx, fy = dfr.flux_xy(k_wall)
    n_out = 1 if dfr.N == 0 else ned
    out = []
    for j, e in enumerate(wall):
        en = int(t.edgeNumL[e])
        offset = en * ned + 2 * ni
        offset2 = en * (ned + 1) + 1                 # skips the vertex node
        for i in range(n_out):
            row = "%.5e, %.5e," % (fx[offset + i, j], fy[offset + i, j])
            for f in fields:
                row += " %.5e," % f[j, i + offset2]
            out.append(row + "\n")
    return "".join(out), (len(wall), n_out)


def output_final(c, q, mesh_file="", outdir=".", out=print):
    """Write the files of OutputFinal for `c.Case`; returns the list of paths written."""
    from gocfd_b200.host.euler2d import FREESTREAM, SHOCKTUBE
    written = []
    if c.Case == SHOCKTUBE:
        st = SODShockTube(4 * c.DFR.K // 5, c.DFR)          # euler.go:771
        st.interpolate_fields(q)
        num, ana = shocktube_files(st, mesh_file, c.FinalTime)
        for name, text in (("shocktube.dat", num), ("shocktube_analytic.dat", ana)):
            path = os.path.join(outdir, name)
            with open(path, "w") as f:
                f.write(text)
            written.append(path)
    elif c.Case == FREESTREAM and len(c.ip.PlotFields) != 0:
        text, (n_wall, n_pts) = wall_plot_data(c, q)
        path = os.path.join(outdir, "plotfile.dat")
        with open(path, "w") as f:
            f.write(text)
        written.append(path)
        out("Output plot data for wall, dimensions: %d Wall edges by %d points each" % (n_wall, n_pts))
    return written
