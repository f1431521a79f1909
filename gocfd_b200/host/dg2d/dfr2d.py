"""DFR2D start-up: reference elements, edge table, geometry.  Host side, runs once.

This is the host-side producer of every constant array that crosses the C ABI
(SURVEY.md appendix C).  In a gocfd deployment these arrays come from `*DG2D.DFR2D`; this
module rebuilds them the same way so the device library can be driven and tested where no
Go toolchain exists.  Everything is vectorised so the 8M-triangle scaling mesh builds in
seconds.

Reference: DG2D/dfr_startup.go:43-175, :256-314 (NewDFR2D, GetHk, CutoffFilter2D,
GetEdgeLengths, ProcessGeometry, CalculateJacobian, CalculateFaceNorms),
DG2D/triangulation.go:19-198, :230-235 (edge table, owner = first registering element,
IInII, outward normals), types/elemental.go:12-49 (EdgeKey packing),
DG2D/dfr_shock_capturing.go:43-140 (sensor matrices, modal filter).
"""
import math

import numpy as np

from .elements import LagrangeElement2D, RTElement
from .. import readfiles as rf


def edge_key(a, b):
    """uint64 = min(v) + max(v) << 32 (types/elemental.go:14-33)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return np.minimum(a, b) + (np.maximum(a, b) << np.uint64(32))


class EdgeTable:
    """Flattened edge map.  Edge index order = order of the owner's half-edge (k*3 + e).

    Per edge: kL/kR (owner / neighbour element, kR = -1 for a boundary edge), local edge
    numbers, number of connected triangles, BC flag, EdgeKey.  Per (element, local edge):
    index into this table and an owner bit.
    """

    def __init__(self, etov, nverts):
        k = etov.shape[0]
        va = etov.reshape(-1)                                  # half-edge h = 3k+e starts at va
        vb = np.roll(etov, -1, axis=1).reshape(-1)             # ... and ends at vb
        keys = edge_key(va, vb)
        order = np.argsort(keys, kind="stable")                # groups share a key; h ascending inside
        sk = keys[order]
        first = np.ones(sk.shape[0], dtype=bool)
        first[1:] = sk[1:] != sk[:-1]
        group_start = np.flatnonzero(first)
        group_size = np.diff(np.append(group_start, sk.shape[0]))
        if group_size.max(initial=0) > 2:
            raise ValueError("incorrect edge construction, more than two connected triangles")
        owner_h = order[group_start]
        second_h = np.where(group_size == 2, order[np.minimum(group_start + 1, sk.shape[0] - 1)], -1)
        # canonical order: by the owner's half-edge index
        canon = np.argsort(owner_h, kind="stable")
        owner_h, second_h = owner_h[canon], second_h[canon]
        self.NE = owner_h.shape[0]
        self.key = keys[owner_h]
        self.kL = (owner_h // 3).astype(np.int32)
        self.edgeNumL = (owner_h % 3).astype(np.int32)
        self.kR = np.where(second_h >= 0, second_h // 3, -1).astype(np.int32)
        self.edgeNumR = np.where(second_h >= 0, second_h % 3, 0).astype(np.int32)
        self.nConn = np.where(second_h >= 0, 2, 1).astype(np.int32)
        self.bcType = np.zeros(self.NE, dtype=np.int32)
        # element -> edge index, owner bit
        self.EtoEdge = np.empty(3 * k, dtype=np.int32)
        self.EtoEdge[owner_h] = np.arange(self.NE, dtype=np.int32)
        has2 = second_h >= 0
        self.EtoEdge[second_h[has2]] = np.arange(self.NE, dtype=np.int32)[has2]
        self.EtoEdge = self.EtoEdge.reshape(k, 3)
        self.isOwner = np.zeros(3 * k, dtype=bool)
        self.isOwner[owner_h] = True
        self.isOwner = self.isOwner.reshape(k, 3)
        # neighbour element across each local edge, -1 at boundaries (Triangulation.EtoE)
        self.EtoE = np.full(3 * k, -1, dtype=np.int32)
        self.EtoE[owner_h[has2]] = self.kR[has2]
        self.EtoE[second_h[has2]] = self.kL[has2]
        self.EtoE = self.EtoE.reshape(k, 3)
        self._sorted_keys = None

    def lookup(self, a, b):
        """Edge indices of vertex pairs (a, b)."""
        if self._sorted_keys is None:
            self._key_order = np.argsort(self.key, kind="stable")
            self._sorted_keys = self.key[self._key_order]
        q = edge_key(a, b)
        pos = np.searchsorted(self._sorted_keys, q)
        pos = np.minimum(pos, self.NE - 1)
        if not np.all(self._sorted_keys[pos] == q):
            raise KeyError("boundary edge not present in the triangulation")
        return self._key_order[pos]

    def apply_bcs(self, bc_edges):
        for tag, pairs in bc_edges.items():
            flag = rf.bc_flag_from_tag(tag)
            if flag in (rf.BC_Periodic, rf.BC_PeriodicReversed):
                # The reference's periodic pairing depends on Go map iteration order
                # (DG2D/dfr_startup.go:296-310 with Euler2D/edges.go:436-438); out of scope.
                raise NotImplementedError("periodic boundaries are not supported (see DESIGN.md)")
            if flag not in (rf.BC_Far, rf.BC_IVortex, rf.BC_Wall, rf.BC_In, rf.BC_Out, rf.BC_Cyl):
                raise NotImplementedError("BC type of tag [%s] not implemented yet" % tag)
            if len(pairs):
                self.bcType[self.lookup(pairs[:, 0], pairs[:, 1])] = flag


class ShockFinderMatrices:
    """Persson modal sensor operators: Clipper = V cut Vinv, D = I - Clipper, P = M D, ModeFilter."""

    def __init__(self, el):
        n, np_ = el.N, el.Np
        order = np.array(el.JB2D.OrderAtJ)
        cut = np.where(order >= n, 0.0, 1.0)
        self.Clipper = el.JB2D.V @ np.diag(cut) @ el.JB2D.Vinv
        self.D = np.eye(np_) - self.Clipper
        self.MassMatrix = el.MassMatrix
        self.P = self.MassMatrix @ self.D
        alpha, s = recommended_filter_parameters(n)
        mf = np.ones(np_)
        if n > 1:
            nz = order > 0
            mf[nz] = np.exp(-alpha * np.power(order[nz] / float(n), float(s)))
        self.ModeFilter = mf


def recommended_filter_parameters(p):
    if p <= 2:
        return 4.0, 4
    table = {3: (6.0, 4), 4: (8.0, 5), 5: (10.0, 6), 6: (14.0, 6), 7: (18.0, 7), 8: (24.0, 8)}
    if p in table:
        return table[p]
    return 30.0 + 2.0 * (p - 9), 8 + (p - 9) // 2


class DFR2D:
    def __init__(self, n, mesh=None):
        if n < 0:
            raise ValueError("Polynomial order must be >= 0, have %d" % n)
        self.N = n
        self.SolutionElement = LagrangeElement2D(n)
        self.FluxElement = RTElement(n + 1)
        rt = self.FluxElement
        r_edge, s_edge = rt.edge_locations(rt.R), rt.edge_locations(rt.S)
        self.FluxEdgeInterp = self.SolutionElement.JB2D.interp_matrix(r_edge, s_edge)
        self.FluxDr, self.FluxDs = self.SolutionElement.derivative_matrices(rt.R, rt.S)
        if mesh is not None:
            self.process_geometry(mesh)

    # ---- geometry -------------------------------------------------------------------
    def process_geometry(self, mesh):
        self.mesh = mesh
        self.K = mesh.K
        self.VX, self.VY = mesh.VX, mesh.VY
        etov = mesh.EToV
        self.EToV = etov
        self.Tris = EdgeTable(etov, len(mesh.VX))
        self.Tris.apply_bcs(mesh.BCEdges)
        vx, vy = self.VX, self.VY
        v1x, v2x, v3x = vx[etov[:, 0]], vx[etov[:, 1]], vx[etov[:, 2]]
        v1y, v2y, v3y = vy[etov[:, 0]], vy[etov[:, 1]], vy[etov[:, 2]]
        xr, yr = 0.5 * (v2x - v1x), 0.5 * (v2y - v1y)
        xs, ys = 0.5 * (v3x - v1x), 0.5 * (v3y - v1y)
        self.J = np.stack([xr, xs, yr, ys], axis=1)
        self.Jdet = xr * ys - xs * yr
        self.Jinv = np.stack([ys, -xs, -yr, xr], axis=1) / self.Jdet[:, None]
        # Face normals / IInII per (local edge, element), index [edge, k] == k + K*edge
        ax = np.stack([v1x, v2x, v3x])           # start vertex of local edge e
        ay = np.stack([v1y, v2y, v3y])
        bx = np.stack([v2x, v3x, v1x])           # end vertex
        by = np.stack([v2y, v3y, v1y])
        dx, dy = ax - bx, ay - by                # x2 - x1 with (x1, x2) = (end, start)
        nrm = np.sqrt(dx * dx + dy * dy)
        oonorm = 1.0 / nrm
        self.FaceNorm = np.stack([-dy * oonorm, dx * oonorm])          # [2, 3, K]
        edge_norm = np.sqrt((-dy) * (-dy) + dx * dx)
        self.IInII = np.stack([edge_norm[0] / 2.0,
                               edge_norm[1] / (2.0 * math.sqrt(2.0)),
                               edge_norm[2] / 2.0])                    # [3, K]
        self.EdgeLenMax = self.Jdet / self.IInII.max(axis=0)
        perimeter = self.IInII[0] + self.IInII[1] + self.IInII[2]
        self.EdgeLenMinR = 4 * 2 * self.Jdet / perimeter

    def hk(self):
        return self.EdgeLenMax / float((self.N + 1) * (self.N + 1))

    def edge_length(self):
        """Physical length of every edge from the owner's IInII (Edge.GetEdgeLength)."""
        t = self.Tris
        iin = self.IInII[t.edgeNumL, t.kL]
        return np.where(t.edgeNumL == 1, (2.0 * math.sqrt(2.0)) * iin, 2.0 * iin)

    def local_coords(self, r, s, elems=None):
        """Physical (X, Y) [len(r), K'] of reference points (CalculateElementLocalGeometry)."""
        etov = self.EToV if elems is None else self.EToV[elems]
        r = np.asarray(r)[:, None]
        s = np.asarray(s)[:, None]
        out = []
        for v in (self.VX, self.VY):
            a, b, c = v[etov[:, 0]][None, :], v[etov[:, 1]][None, :], v[etov[:, 2]][None, :]
            out.append((((r + s) * -1.0) * a + (r + 1.0) * b + (s + 1.0) * c) * 0.5)
        return out[0], out[1]

    def solution_xy(self, elems=None):
        el = self.SolutionElement
        return self.local_coords(el.R, el.S, elems)

    def flux_xy(self, elems=None):
        rt = self.FluxElement
        return self.local_coords(rt.R, rt.S, elems)

    def metrics(self):
        """DXMetric, DYMetric [NpFlux, K] (CalculateRTBasedDerivativeMetrics); small meshes only."""
        rt = self.FluxElement
        ni, ne, k = rt.NpInt, rt.NpEdge, self.K
        dxm = np.empty((rt.Np, k))
        dym = np.empty((rt.Np, k))
        dxm[:ni], dxm[ni:2 * ni] = self.Jinv[:, 0], self.Jinv[:, 2]
        dym[:ni], dym[ni:2 * ni] = self.Jinv[:, 1], self.Jinv[:, 3]
        oojd = 1.0 / self.Jdet
        for fn in range(3):
            rows = slice(2 * ni + fn * ne, 2 * ni + (fn + 1) * ne)
            dxm[rows] = oojd * self.FaceNorm[0, fn] * self.IInII[fn]
            dym[rows] = oojd * self.FaceNorm[1, fn] * self.IInII[fn]
        return dxm, dym

    def shock_finder(self):
        return ShockFinderMatrices(self.SolutionElement)

    def graph_rs(self):
        """GetRSForGraphMesh (DG2D/graphics_support.go:130-162): per edge the leading vertex then the RT edge
        points, counter-clockwise, followed by the interior points -- 3(1+NpEdge)+NpInt graph nodes."""
        rt = self.FluxElement
        ni, ne = rt.NpInt, rt.NpEdge
        vr, vs = [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]
        r, s = [], []
        for n in range(3):
            r.append(vr[n]); s.append(vs[n])
            off = 2 * ni + n * ne
            r.extend(rt.R[off:off + ne]); s.extend(rt.S[off:off + ne])
        r.extend(rt.R[:ni]); s.extend(rt.S[:ni])
        return np.array(r), np.array(s)

    def graph_interp(self):
        """DFR.GraphInterp = SolutionBasis.GetInterpMatrix(GraphR, GraphS) (DG2D/dfr_startup.go:62-63)."""
        r, s = self.graph_rs()
        return np.ascontiguousarray(self.SolutionElement.JB2D.interp_matrix(r, s))

    def barycentric_coords(self):
        """[NpFlux, 3] vertex interpolation weights at every RT point (dissipation.go:414-449)."""
        rt = self.FluxElement
        rr = np.array([[1.0, 1.0, 1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]])
        rrinv = np.linalg.inv(rr)
        c = np.stack([np.ones(rt.Np), rt.R, rt.S])
        return (rrinv @ c).T.copy()


def new_dfr2d(n, mesh_file=None):
    mesh = rf.read_mesh(mesh_file) if mesh_file else None
    return DFR2D(n, mesh)
