"""Partition-local problems: the slice of a Problem that dfr2d_create_window needs for one
partition -- its own PartitionMap range plus the ring of elements that share an edge (with the limiter: a vertex) with
it -- so that a process driving one GPU never holds the global arrays of an 8M-element run.

`window_problem` cuts the slice out of a global Problem (any mesh whose numbering is spatially compact, e.g. after
dfr2d_rcm_order; tests); `structured_window_case` builds it directly for the synthetic strip meshes of bench.py without
ever forming the global mesh.
"""
import copy

import numpy as np


def split_1d(max_index, n_parts, n):
    """utils.PartitionMap.Split1D (utils/parallel_utils.go:172-192)."""
    npart, rem = divmod(max_index, n_parts)
    if rem != 0:
        start_add, end_add = (rem, 0) if n + 1 > rem else (n, 1)
    else:
        start_add = end_add = 0
    lo = n * npart + start_add
    return lo, lo + npart + end_add


def window_range(p, n_parts, part):
    """[w0, w1): the smallest contiguous element range holding the partition and its edge (vertex) ring."""
    k0, k1 = split_1d(p.K, n_parts, part)
    kl, kr = np.asarray(p.edge_kL, dtype=np.int64), np.asarray(p.edge_kR, dtype=np.int64)
    two = np.asarray(p.edge_nconn) == 2
    lm = (kl >= k0) & (kl < k1)
    rm = two & (kr >= k0) & (kr < k1)
    need = [np.array([k0, k1 - 1]), kr[two & lm], kl[rm]]
    if p.Dissipation:
        etov = np.asarray(p.EToV)
        mask = np.zeros(p.NV, dtype=bool)
        mask[etov[k0:k1].ravel()] = True
        need.append(np.flatnonzero(mask[etov].any(axis=1)))
    need = np.concatenate(need)
    return int(need.min()), int(need.max()) + 1


def window_problem(p, n_parts, part):
    """(window Problem, (K_global, k_offset)) for dfr2d_create_window / lib.Dfr2d(..., window=...)."""
    w0, w1 = window_range(p, n_parts, part)
    kw = w1 - w0
    kl, kr = np.asarray(p.edge_kL, dtype=np.int64), np.asarray(p.edge_kR, dtype=np.int64)
    nconn = np.asarray(p.edge_nconn)
    two = nconn == 2
    keep = (kl >= w0) & (kl < w1) & (~two | ((kr >= w0) & (kr < w1)))
    new_index = np.cumsum(keep) - 1
    q = copy.copy(p)
    q.K, q.NE = kw, int(keep.sum())
    q.Jdet = np.ascontiguousarray(p.Jdet[w0:w1])
    q.Jinv = np.ascontiguousarray(np.asarray(p.Jinv).reshape(p.K, 4)[w0:w1])
    for name in ("FaceNormX", "FaceNormY", "IInII"):
        setattr(q, name, np.ascontiguousarray(np.asarray(getattr(p, name)).reshape(3, p.K)[:, w0:w1]).ravel())
    q.EdgeLenMax = np.ascontiguousarray(p.EdgeLenMax[w0:w1])
    q.EToV = np.ascontiguousarray(np.asarray(p.EToV)[w0:w1])                     # vertex ids stay global
    q.edge_kL = (kl[keep] - w0).astype(np.int32)
    q.edge_kR = np.where(two[keep], kr[keep] - w0, kr[keep]).astype(np.int32)
    for name in ("edge_numL", "edge_numR", "edge_nconn", "edge_bc", "edge_len"):
        setattr(q, name, np.ascontiguousarray(np.asarray(getattr(p, name))[keep]))
    ete = np.asarray(p.EtoEdge).reshape(p.K, 3)[w0:w1]
    q.EtoEdge = np.where(keep[ete], new_index[ete], 0).astype(np.int32)         # ghost elements' outer edges: unused
    bpe = np.asarray(p.bp_edge, dtype=np.int64)
    bkeep = keep[bpe] if len(bpe) else np.zeros(0, dtype=bool)
    q.NBP = int(bkeep.sum())
    q.bp_edge = new_index[bpe[bkeep]].astype(np.int32) if q.NBP else np.zeros(0, dtype=np.int32)
    ne = p.NpEdge
    q.bp_x = np.ascontiguousarray(np.asarray(p.bp_x).reshape(-1, ne)[bkeep]) if q.NBP else np.zeros((0, ne))
    q.bp_y = np.ascontiguousarray(np.asarray(p.bp_y).reshape(-1, ne)[bkeep]) if q.NBP else np.zeros((0, ne))
    return q, (int(p.K), w0)


def structured_window_case(build_euler, nx, ny, n_parts, part, x0, x1, y0, y1, **mesh_kw):
    """The window of partition `part` of the nx x ny x 2 strip mesh of host.meshgen, built from the rows it needs only.
    build_euler(mesh) -> Euler.  Returns (Euler of the window mesh, (K_global, k_offset)).

    Row-major numbering: element k lies in row k // (2 nx); the ring of a contiguous element range is at most one row
    below and one above (edge and vertex neighbours alike).  The window mesh takes its vertex coordinates from the SAME
    linspace arrays as the global mesh, so every metric is bitwise the global one; only true domain boundaries carry a
    tag, the cut sides stay untagged (their edges belong to ghost elements and are never evaluated)."""
    from .readfiles import Mesh2D
    k_global = 2 * nx * ny
    k0, k1 = split_1d(k_global, n_parts, part)
    j0 = max(0, k0 // (2 * nx) - 1)
    j1 = min(ny, (k1 - 1) // (2 * nx) + 2)                   # rows [j0, j1)
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)[j0:j1 + 1]
    nyw = j1 - j0
    vx = np.tile(xs, nyw + 1)
    vy = np.repeat(ys, nx + 1)
    i = np.arange(nx)[None, :]
    j = np.arange(nyw)[:, None]
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10, v01 = v00 + 1, v00 + (nx + 1)
    v11 = v01 + 1
    etov = np.empty((nx * nyw, 2, 3), dtype=np.int64)
    etov[:, 0] = np.stack([v00, v10, v11], axis=1)
    etov[:, 1] = np.stack([v00, v11, v01], axis=1)
    etov = etov.reshape(-1, 3)
    tag = mesh_kw.get("tag", "wall")
    side_tags = mesh_kw.get("side_tags") or {}
    b = np.arange(nx)
    s = np.arange(nyw)
    sides = {"right": np.stack([s * (nx + 1) + nx, (s + 1) * (nx + 1) + nx], axis=1),
             "left": np.stack([(s + 1) * (nx + 1), s * (nx + 1)], axis=1)}
    if j0 == 0:
        sides["bottom"] = np.stack([b, b + 1], axis=1)
    if j1 == ny:
        sides["top"] = np.stack([nyw * (nx + 1) + b + 1, nyw * (nx + 1) + b], axis=1)
    groups = {}
    for name, edges in sides.items():
        groups.setdefault(side_tags.get(name, tag), []).append(edges)
    mesh = Mesh2D(vx, vy, etov, {t: np.concatenate(e) for t, e in groups.items()})
    c = build_euler(mesh)
    # vertex ids must be global for the shared-vertex exchange: shift the window's row-major vertex numbering
    c.problem.EToV = np.ascontiguousarray(c.problem.EToV + j0 * (nx + 1)).astype(np.int32)
    c.problem.NV = (nx + 1) * (ny + 1)
    return c, (k_global, 2 * nx * j0)
