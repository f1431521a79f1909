"""2D mesh readers (SU2 and Gambit neutral) and an SU2 writer.  Host side only.

Output convention matches the reference readers: K, VX, VY, EToV (K x 3, zero-based vertex
ids in the file's element order) and a BC map {tag(lower-case) -> list of directed vertex
pairs}.  The tag's base name (text before the first '-') selects the BC type.

Reference: readfiles/readSU2Grid.go:27-177 (grammar), readfiles/readGambitGrid.go:21-166,
:334-363, types/cfd.go:13-75 (BC flag numbering and tag parsing).
"""
import numpy as np

# BC flag numbering is part of the device ABI (types/cfd.go:13-26).
BC_None, BC_In, BC_Dirichlet, BC_Slip, BC_Far, BC_Wall, BC_Cyl, BC_Neuman, BC_Out, \
    BC_IVortex, BC_Periodic, BC_PeriodicReversed = range(12)

BC_NAME_MAP = {
    "inflow": BC_In, "in": BC_In, "out": BC_Out, "outflow": BC_Out, "wall": BC_Wall,
    "far": BC_Far, "cyl": BC_Cyl, "dirichlet": BC_Dirichlet, "neuman": BC_Neuman,
    "slip": BC_Slip, "periodic": BC_Periodic, "periodicreversed": BC_PeriodicReversed,
}


def bc_flag_from_tag(tag):
    base = tag.strip().lower()
    ind = base.find("-")
    if ind > 0:
        base = base[:ind]
    if base not in BC_NAME_MAP:
        raise KeyError("unable to find BC with base name: [%s], full tag: [%s]" % (base, tag))
    return BC_NAME_MAP[base]


class Mesh2D:
    def __init__(self, vx, vy, etov, bc_edges):
        self.VX = np.ascontiguousarray(vx, dtype=np.float64)
        self.VY = np.ascontiguousarray(vy, dtype=np.float64)
        self.EToV = np.ascontiguousarray(etov, dtype=np.int64).reshape(-1, 3)
        self.K = self.EToV.shape[0]
        # {tag: (n, 2) int array of vertex pairs}
        self.BCEdges = {k.strip().lower(): np.asarray(v, dtype=np.int64).reshape(-1, 2)
                        for k, v in bc_edges.items()}


def renumber_elements(mesh, order):
    """Mesh with element `new` = old element order[new] (vertices and boundary edges untouched)."""
    order = np.asarray(order, dtype=np.int64)
    assert sorted(order.tolist()) == list(range(mesh.K))
    return Mesh2D(mesh.VX, mesh.VY, mesh.EToV[order], dict(mesh.BCEdges))


def _su2_tokens(path):
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("%"):
                continue
            yield line


def read_su2(path):
    lines = _su2_tokens(path)

    def number(line, key):
        if "=" not in line:
            raise ValueError("badly formed input line [%s], should have an =" % line)
        return int(line.split("=", 1)[1].split()[0])

    ndime = number(next(lines), "NDIME")
    if ndime != 2:
        raise ValueError("only 2D SU2 meshes are supported, have NDIME=%d" % ndime)
    k = number(next(lines), "NELEM")
    etov = np.empty((k, 3), dtype=np.int64)
    for i in range(k):
        parts = next(lines).split()
        if int(parts[0]) != 5:
            raise ValueError("unable to deal with non-triangular elements right now")
        etov[i] = (int(parts[1]), int(parts[2]), int(parts[3]))
    nv = number(next(lines), "NPOIN")
    vx = np.empty(nv)
    vy = np.empty(nv)
    for i in range(nv):
        parts = next(lines).split()
        vx[i], vy[i] = float(parts[0]), float(parts[1])
    nbc = number(next(lines), "NMARK")
    bcs = {}
    for _ in range(nbc):
        tagline = next(lines)
        tag = tagline.split("=", 1)[1].strip().lower()
        ne = number(next(lines), "MARKER_ELEMS")
        edges = np.empty((ne, 2), dtype=np.int64)
        for i in range(ne):
            parts = next(lines).split()
            if int(parts[0]) != 3:
                raise ValueError("BCs should only contain line elements in 2D")
            edges[i] = (int(parts[1]), int(parts[2]))
        # duplicate tags append to a common list (types/cfd.go:91-109)
        bcs[tag] = np.concatenate([bcs[tag], edges]) if tag in bcs else edges
    return Mesh2D(vx, vy, etov, bcs)


def write_su2(path, mesh):
    with open(path, "w") as f:
        f.write("NDIME= 2\nNELEM= %d\n" % mesh.K)
        for i, (a, b, c) in enumerate(mesh.EToV):
            f.write("5\t%d\t%d\t%d\t%d\n" % (a, b, c, i))
        f.write("NPOIN= %d\n" % len(mesh.VX))
        for i, (x, y) in enumerate(zip(mesh.VX, mesh.VY)):
            f.write("\t%.17e\t%.17e\t%d\n" % (x, y, i))
        f.write("NMARK= %d\n" % len(mesh.BCEdges))
        for tag, edges in mesh.BCEdges.items():
            f.write("MARKER_TAG= %s\nMARKER_ELEMS= %d\n" % (tag, len(edges)))
            for a, b in edges:
                f.write("3\t%d\t%d\n" % (a, b))


def read_gambit_2d(path):
    with open(path, "r") as f:
        lines = f.read().splitlines()
    pos = 6
    hdr = lines[pos].split()
    nv, k, nmats, nbcs, nsd = (int(x) for x in hdr[:5])
    if nsd != 2:
        raise ValueError("space dimensions not 2")
    pos += 3
    vx = np.empty(nv)
    vy = np.empty(nv)
    for i in range(nv):
        parts = lines[pos + i].split()
        vx[i], vy[i] = float(parts[1]), float(parts[2])
    pos += nv + 2
    etov = np.empty((k, 3), dtype=np.int64)
    for i in range(k):
        parts = lines[pos + i].split()
        etov[int(parts[0]) - 1] = (int(parts[3]) - 1, int(parts[4]) - 1, int(parts[5]) - 1)
    pos += k + 2
    for _ in range(nmats):
        parts = lines[pos].replace(":", ": ").split()
        elnum = int(parts[parts.index("ELEMENTS:") + 1])
        pos += 3
        pos += elnum // 10 + (1 if elnum % 10 else 0)
        pos += 2
    bcs = {}
    for ib in range(nbcs):
        if ib != 0:
            pos += 1
        parts = lines[pos].split()
        tag = parts[0].strip().lower()
        numfaces = int(parts[2])
        pos += 1
        edges = np.empty((numfaces, 2), dtype=np.int64)
        for i in range(numfaces):
            kp1, _, face = (int(x) for x in lines[pos + i].split()[:3])
            v = etov[kp1 - 1]
            edges[i] = {1: (v[0], v[1]), 2: (v[1], v[2]), 3: (v[2], v[0])}[face]
        pos += numfaces + 1
        bcs[tag] = np.concatenate([bcs[tag], edges]) if tag in bcs else edges
    return Mesh2D(vx, vy, etov, bcs)


def read_mesh(path):
    p = path.strip()
    if len(p) < 4 or p[-4] != ".":
        raise ValueError("unable to determine file type from name: %s" % path)
    ext = p[-3:]
    if ext == "neu":
        return read_gambit_2d(p)
    if ext == "su2":
        return read_su2(p)
    raise ValueError("unsupported file type: %s" % path)
