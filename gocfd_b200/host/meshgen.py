"""Synthetic structured triangle meshes for the benchmark configs.

nx x ny squares on [x0,x1] x [y0,y1], each split into two counter-clockwise right triangles,
row-major numbering so that contiguous element ranges (utils.PartitionMap) are horizontal strips.
All four sides carry one boundary tag ("wall" becomes BC_IVortex for the IVORTEX case,
euler.go:781-787).  Same grammar as the SU2 files the reference reads (readSU2Grid.go:27-177);
`readfiles.write_su2` serialises it for a Go-side run.
"""
import numpy as np

from .readfiles import Mesh2D


def structured_tri_mesh(nx, ny, x0=-10.0, x1=10.0, y0=-10.0, y1=10.0, tag="wall", side_tags=None):
    """side_tags = {"bottom": .., "right": .., "top": .., "left": ..} overrides `tag` per side (e.g. the shock-tube
    layout of test_cases/Euler2D/shock-tube: left "in", right "out", top/bottom "wall")."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    vx = np.tile(xs, ny + 1)
    vy = np.repeat(ys, nx + 1)
    i = np.arange(nx)[None, :]
    j = np.arange(ny)[:, None]
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    etov = np.empty((nx * ny, 2, 3), dtype=np.int64)
    etov[:, 0] = np.stack([v00, v10, v11], axis=1)      # lower-right triangle
    etov[:, 1] = np.stack([v00, v11, v01], axis=1)      # upper-left triangle
    etov = etov.reshape(-1, 3)
    b = np.arange(nx)
    bottom = np.stack([b, b + 1], axis=1)
    top = np.stack([ny * (nx + 1) + b + 1, ny * (nx + 1) + b], axis=1)
    s = np.arange(ny)
    right = np.stack([s * (nx + 1) + nx, (s + 1) * (nx + 1) + nx], axis=1)
    left = np.stack([(s + 1) * (nx + 1), s * (nx + 1)], axis=1)
    if side_tags is None:
        return Mesh2D(vx, vy, etov, {tag: np.concatenate([bottom, right, top, left])})
    groups = {}
    for name, edges in (("bottom", bottom), ("right", right), ("top", top), ("left", left)):
        groups.setdefault(side_tags.get(name, tag), []).append(edges)
    return Mesh2D(vx, vy, etov, {t: np.concatenate(e) for t, e in groups.items()})
