"""CPU-only tests of the boundary: the C-ABI library loads and exports every symbol the header
declares, and the multi-partition bookkeeping (partition ranges, local edge tables, ghost columns,
halo message lists) is bit-exact against the reference's PartitionMap / edge ownership rules.
No compute call is made here (there is no GPU in the CPU test environment)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, mesh_path
from gocfd_b200.host.euler2d import Euler, PartitionMap
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh


def _case(n, mesh, **kw):
    base = dict(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=n, FinalTime=1.0, MaxIterations=10,
                Gamma=1.4, Minf=0.1)
    base.update(kw)
    return Euler(InputParameters2D(**base), mesh)


def test_library_exports_every_declared_symbol():
    from gocfd_b200 import lib
    header = open(os.path.join(ROOT, "include", "dfr2d.h")).read()
    declared = set(re.findall(r"\b(dfr2d_[a-z_0-9]+)\s*\(", header))
    assert declared == set(lib.EXPORTS)
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name
    # no torch / C++ types leak through the boundary: every exported dfr2d_ symbol is unmangled C
    assert not re.search(r"_Z\w*dfr2d_(create|step|set_state)", out)


def test_create_rejects_bad_requests_without_touching_cuda():
    from gocfd_b200 import lib
    c = _case(2, structured_tri_mesh(4, 4))
    with pytest.raises(lib.Dfr2dError):
        lib.Dfr2d(c.problem, n_parts=2, part=2)
    c.problem.N = 9
    with pytest.raises(lib.Dfr2dError):
        lib.Dfr2d(c.problem)


def _shuffled_grid(seed):
    """A structured grid whose elements are renumbered at random: contiguous Split1D ranges are then scattered in
    space, so almost every edge is cut and vertices are shared by many partitions (the worst case for the halo lists)."""
    from gocfd_b200.host import readfiles as rf
    m = structured_tri_mesh(7, 5)
    return rf.renumber_elements(m, np.random.default_rng(seed).permutation(m.K))


@pytest.mark.parametrize("mesh_name,n_parts", [("grid", 2), ("grid", 3), ("grid", 8), ("naca", 2), ("naca", 5),
                                               ("shuffled", 2), ("shuffled", 5), ("shuffled", 9)])
def test_partition_plan_tables(mesh_name, n_parts):
    from gocfd_b200 import lib
    mesh = {"grid": lambda: structured_tri_mesh(12, 10), "naca": lambda: mesh_path("mesh_NACA0012_inv.su2"),
            "shuffled": lambda: _shuffled_grid(n_parts)}[mesh_name]()
    c = _case(1, mesh, InitType="Freestream" if mesh_name == "naca" else "IVortex")
    p = c.problem
    pm = PartitionMap(n_parts, p.K)
    plans = [lib.Plan(p, n_parts, r) for r in range(n_parts)]
    ne = p.NpEdge
    owned_edges = 0
    for r, pl in enumerate(plans):
        assert (pl.k0, pl.k1) == pm.get_bucket_range(r)             # Split1D ranges, bit-exact
        mine = lambda k: pl.k0 <= k < pl.k1                         # noqa: E731
        ge = pl.global_edge
        # local edge set = every edge touching an owned element
        want = [e for e in range(p.NE) if mine(p.edge_kL[e]) or (p.edge_nconn[e] == 2 and mine(p.edge_kR[e]))]
        assert sorted(ge.tolist()) == want
        col2glob = np.concatenate([np.arange(pl.k0, pl.k1), pl.ghost_global])
        assert np.array_equal(col2glob[pl.kL], p.edge_kL[ge])
        shared = p.edge_nconn[ge] == 2
        assert np.array_equal(col2glob[pl.kR[shared]], p.edge_kR[ge][shared])
        assert np.all(pl.kR[~shared] < 0)
        assert np.array_equal(pl.meta & 3, p.edge_numL[ge])
        assert np.array_equal((pl.meta >> 2) & 3, p.edge_numR[ge])
        assert np.array_equal((pl.meta >> 4) & 15, p.edge_bc[ge])
        # element -> edge slot with the owner bit (replaces both Go maps, edges.go:101-111)
        for le in range(3):
            s = pl.etoe[le, :pl.K]
            slot = np.where(s >= 0, s, -1 - s)
            gk = np.arange(pl.k0, pl.k1)
            assert np.array_equal(ge[slot], p.EtoEdge[gk, le])
            assert np.array_equal(s >= 0, p.edge_kL[p.EtoEdge[gk, le]] == gk)
        owned_edges += int(np.sum([mine(k) for k in p.edge_kL[ge]]))
    assert owned_edges == p.NE                                      # PartitionEdgesByK: every edge has one owner
    # halo lists: what r sends to s is what s expects from r, edge by edge
    rng = np.random.default_rng(0)
    qface = rng.standard_normal((4, 3 * ne, p.K))                   # a global Q_Face
    local = []
    for pl in plans:
        q = np.zeros((4, 3 * ne, pl.Kp))
        q[:, :, :pl.K] = qface[:, :, pl.k0:pl.k1]
        local.append(q)
    per = 4 * ne
    sent = {}
    for r, pl in enumerate(plans):
        buf = np.empty(pl.n_cut * per)
        for ci in range(pl.n_cut):
            rows = pl.send_row0[ci] + np.arange(ne)
            buf[ci * per:(ci + 1) * per] = local[r][:, rows, pl.send_elem[ci]].reshape(-1)
        off = np.concatenate([[0], np.cumsum(pl.send_counts)])
        for s in range(n_parts):
            sent[(r, s)] = buf[off[s]:off[s + 1]]
        assert pl.send_counts[r] == 0
    for s, pl in enumerate(plans):
        assert [len(sent[(r, s)]) for r in range(n_parts)] == pl.recv_counts.tolist()
        buf = np.concatenate([sent[(r, s)] for r in range(n_parts)])
        for ci in range(pl.n_cut):
            rows = pl.recv_row0[ci] + np.arange(ne)
            local[s][:, rows, pl.recv_col[ci]] = buf[ci * per:(ci + 1) * per].reshape(4, ne)
    # after the exchange every local edge sees exactly the global values on both sides
    for r, pl in enumerate(plans):
        ge = pl.global_edge
        i = np.arange(ne)
        rows_l = p.edge_numL[ge][:, None] * ne + i[None, :]
        for nvar in range(4):
            have = local[r][nvar][rows_l, pl.kL[:, None]]
            want = qface[nvar][rows_l, p.edge_kL[ge][:, None]]
            assert np.array_equal(have, want)
        sh = np.flatnonzero(p.edge_nconn[ge] == 2)
        rows_r = p.edge_numR[ge][sh][:, None] * ne + (ne - 1 - i)[None, :]
        for nvar in range(4):
            have = local[r][nvar][rows_r, pl.kR[sh][:, None]]
            want = qface[nvar][rows_r, p.edge_kR[ge][sh][:, None]]
            assert np.array_equal(have, want)


def test_two_rank_gloo_exchange():
    """world_size-2 gloo run of the host side of the multi-GPU step: plan, pack, all_to_all of the halo,
    unpack, MAX-allreduce of the wave speed -- the same calls bench.py issues over NCCL."""
    script = os.path.join(ROOT, "tests", "gloo_halo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [pr.communicate(timeout=300)[0] for pr in procs]
    for pr, out in zip(procs, outs):
        assert pr.returncode == 0, out
        assert "OK" in out, out


@pytest.mark.parametrize("mesh_name,n_parts", [("sod", 2), ("sod", 3), ("naca", 5), ("naca", 8), ("shuffled", 4),
                                               ("shuffled", 7)])
def test_shared_vertex_plan_reproduces_the_global_vertex_max(mesh_name, n_parts):
    """Dissipation across partitions: the per-peer shared-vertex lists are exactly the pairwise
    intersections of the partitions' vertex sets, both sides agree on the order, and local max + exchange + max equals
    MergeElementScalarToVertices over the whole mesh (euler.go:1048-1065)."""
    from gocfd_b200 import lib
    if mesh_name == "sod":
        c = _case(2, mesh_path("sod-aligned-100pts.su2"), InitType="shocktube", Limiter="persson c0", Kappa=5.0)
    elif mesh_name == "shuffled":       # every vertex shared by several scattered partitions
        c = _case(2, _shuffled_grid(n_parts), InitType="IVortex", Limiter="PerssonC0")
    else:
        c = _case(2, mesh_path("mesh_NACA0012_inv.su2"), InitType="Freestream", Minf=0.8, Limiter="PerssonC0")
    p = c.problem
    assert p.Dissipation
    plans = [lib.Plan(p, n_parts, r) for r in range(n_parts)]
    ne = p.NpEdge
    touched = [set(np.unique(p.EToV[pl.k0:pl.k1]).tolist()) for pl in plans]
    rng = np.random.default_rng(3)
    elem_val = rng.random(p.K)
    want = np.zeros(p.NV)
    np.maximum.at(want, p.EToV.reshape(-1), np.repeat(elem_val, 3))
    local = []
    for pl in plans:
        v = np.zeros(p.NV)
        np.maximum.at(v, p.EToV[pl.k0:pl.k1].reshape(-1), np.repeat(elem_val[pl.k0:pl.k1], 3))
        local.append(v)
    merged = [v.copy() for v in local]
    for r, pl in enumerate(plans):
        # Q_Face message carries 3 extra doubles (vertex epsilon of the sender's element) per cut edge
        if pl.n_cut:
            assert int(pl.send_counts.sum()) == pl.n_cut * (4 * ne + 3)
        off = np.concatenate([[0], np.cumsum(pl.vertex_counts // 2)])
        assert pl.vertex_counts[r] == 0
        for s in range(n_parts):
            ids = pl.vertex_ids[off[s]:off[s + 1]]
            assert ids.tolist() == sorted(touched[r] & touched[s]) if s != r else len(ids) == 0
            # what r sends to s is what s expects from r, in the same order
            offs = np.concatenate([[0], np.cumsum(plans[s].vertex_counts // 2)])
            assert np.array_equal(ids, plans[s].vertex_ids[offs[r]:offs[r + 1]])
            np.maximum.at(merged[s], ids, local[r][ids])
    for r in range(n_parts):
        mine = sorted(touched[r])
        assert np.array_equal(merged[r][mine], want[mine])


def test_two_rank_gloo_dissipation_exchange():
    """world_size-2 gloo run of the dissipation path's shared-vertex exchange."""
    script = os.path.join(ROOT, "tests", "gloo_vertex_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29733", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", mesh_path("sod-aligned-100pts.su2")], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [pr.communicate(timeout=300)[0] for pr in procs]
    for pr, out in zip(procs, outs):
        assert pr.returncode == 0, out
        assert "OK" in out, out


def test_rcm_renumbering_shrinks_the_cuts_of_an_unstructured_mesh():
    """PartitionMap.Split1D cuts the element index range, so the halo size is decided by the mesh
    numbering.  RCM order from the C ABI (host-only) on the shipped NACA0012 mesh, 8 partitions."""
    from gocfd_b200 import lib
    from gocfd_b200.host import readfiles as rf
    mesh = rf.read_mesh(mesh_path("mesh_NACA0012_inv.su2"))
    c = _case(1, mesh, InitType="Freestream")
    p = c.problem
    order = lib.rcm_order(p)
    assert sorted(order.tolist()) == list(range(p.K))
    c2 = _case(1, rf.renumber_elements(mesh, order), InitType="Freestream")
    p2 = c2.problem
    assert p2.K == p.K and p2.NE == p.NE and int((p2.edge_nconn == 1).sum()) == int((p.edge_nconn == 1).sum())

    def cut_edges(prob, n_parts):
        pm = PartitionMap(n_parts, prob.K)
        sh = prob.edge_nconn == 2
        bl = np.array([pm.get_bucket(int(k))[0] for k in prob.edge_kL[sh]])
        br = np.array([pm.get_bucket(int(k))[0] for k in prob.edge_kR[sh]])
        return int((bl != br).sum())

    before, after = cut_edges(p, 8), cut_edges(p2, 8)
    assert after < 0.5 * before, (before, after)
    # bandwidth of the element adjacency: what RCM minimises
    sh = p.edge_nconn == 2
    bw0 = int(np.abs(p.edge_kL[sh].astype(np.int64) - p.edge_kR[sh]).max())
    sh2 = p2.edge_nconn == 2
    bw1 = int(np.abs(p2.edge_kL[sh2].astype(np.int64) - p2.edge_kR[sh2]).max())
    assert bw1 < bw0
    # and the plan of the renumbered mesh agrees: total halo of all partitions = 2 x cut edges
    plans = [lib.Plan(p2, 8, r) for r in range(8)]
    assert sum(pl.n_cut for pl in plans) == 2 * after
    print("cut edges at 8 partitions: %d -> %d, adjacency bandwidth %d -> %d" % (before, after, bw0, bw1))


def test_multi_step_rejects_bad_requests_without_touching_cuda():
    import ctypes as C
    from gocfd_b200 import lib
    L = lib.load()
    assert L.dfr2d_multi_step(None, 2, 1, None) == 1
    arr = (C.c_void_p * 2)(None, None)
    assert L.dfr2d_multi_step(arr, 0, 1, None) == 1
    assert L.dfr2d_multi_step(arr, 33, 1, None) == 1
    assert L.dfr2d_multi_step(arr, 2, 1, None) == 1      # null handles


def test_makefile_builds_with_the_flags_of_the_build_entry():
    """gocfd_b200/csrc/Makefile (the Python-free build a Go tree uses) carries exactly the nvcc flags of
    __graft_entry__.build(): sm_100a only, -lineinfo."""
    import __graft_entry__ as g
    mk = open(os.path.join(ROOT, "gocfd_b200", "csrc", "Makefile")).read()
    flags = re.search(r"^NVCCFLAGS := (.*)$", mk, re.M).group(1).split()
    assert flags == g.NVCC_FLAGS
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_multi_create_argument_validation_and_loud_failure_without_a_device():
    """dfr2d_multi_create (one create for the run): bad arguments are refused, and on a machine without a
    GPU the call fails with the CUDA error and leaves every handle slot NULL -- there is no CPU path behind it."""
    import ctypes as C
    from conftest import _cuda_device_count
    from gocfd_b200 import lib
    l = lib.load()
    c = _case(1, structured_tri_mesh(4, 3))
    s, keep = lib.problem_struct(c.problem)
    hs = (C.c_void_p * 3)()
    assert l.dfr2d_multi_create(C.byref(s), 0, None, hs) == 1
    assert b"bad arguments" in l.dfr2d_last_error(None)
    assert l.dfr2d_multi_create(None, 2, None, hs) == 1
    if _cuda_device_count() == 0:
        for devices in (None, (C.c_int32 * 3)(0, 0, 0)):
            hs = (C.c_void_p * 3)(1, 1, 1)
            assert l.dfr2d_multi_create(C.byref(s), 3, devices, hs) != 0
            assert l.dfr2d_last_error(None)
            if devices is not None:
                assert all(h is None for h in hs)
    l.dfr2d_multi_destroy(None, 0)          # tolerated
    assert l.dfr2d_multi_residual(None, 2, None) == 1
    del keep


def _hilbert_reference(etov, vx, vy):
    """numpy restatement of dfr2d_hilbert_order: rank-normalised centroids, canonical xy -> d of a 65536^2 Hilbert curve."""
    ev = np.asarray(etov).reshape(-1, 3)
    k = ev.shape[0]
    side = 65536

    def rank(c):
        idx = np.lexsort((np.arange(k), c))
        r = np.empty(k, dtype=np.int64)
        r[idx] = np.arange(k, dtype=np.int64) * (side - 1) // max(k - 1, 1)
        return r

    cx = (vx[ev[:, 0]] + vx[ev[:, 1]]) + vx[ev[:, 2]]
    cy = (vy[ev[:, 0]] + vy[ev[:, 1]]) + vy[ev[:, 2]]
    x, y = rank(cx), rank(cy)
    d = np.zeros(k, dtype=np.int64)
    s = side // 2
    while s > 0:
        bx, by = ((x & s) > 0).astype(np.int64), ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * bx) ^ by)
        flip = (by == 0) & (bx == 1)
        x, y = np.where(flip, side - 1 - x, x), np.where(flip, side - 1 - y, y)
        x, y = np.where(by == 0, y, x), np.where(by == 0, x, y)
        s //= 2
    return np.lexsort((np.arange(k), d)).astype(np.int32)


@pytest.mark.parametrize("name", ["mesh_NACA0012_inv.su2", "nacaAirfoil-base.su2"])
def test_hilbert_order_beats_rcm_on_the_airfoil_meshes(name):
    """dfr2d_hilbert_order (host only): bit-exact against its numpy restatement, a permutation, and -- on both airfoil
    meshes the reference ships -- fewer cut edges of the Split1D ranges than RCM at 4 and 8 partitions; the plans of the
    renumbered mesh carry exactly those cuts."""
    from gocfd_b200 import lib
    from gocfd_b200.host import readfiles as rf
    mesh = rf.read_mesh(mesh_path(name))
    c = _case(1, mesh, InitType="Freestream")
    p = c.problem
    vx, vy = np.asarray(c.DFR.VX, dtype=np.float64), np.asarray(c.DFR.VY, dtype=np.float64)
    order = lib.hilbert_order(p.EToV, vx, vy)
    assert sorted(order.tolist()) == list(range(p.K))
    np.testing.assert_array_equal(order, _hilbert_reference(p.EToV, vx, vy))

    def cut_edges(prob, n_parts):
        pm = PartitionMap(n_parts, prob.K)
        sh = prob.edge_nconn == 2
        bounds = np.array([pm.get_bucket_range(b)[1] for b in range(n_parts)])
        bl, br = np.searchsorted(bounds, prob.edge_kL[sh], side="right"), np.searchsorted(bounds, prob.edge_kR[sh], side="right")
        return int((bl != br).sum())

    p_h = _case(1, rf.renumber_elements(mesh, order), InitType="Freestream").problem
    p_r = _case(1, rf.renumber_elements(mesh, lib.rcm_order(p)), InitType="Freestream").problem
    for n_parts in (4, 8):
        orig, rcm, hil = cut_edges(p, n_parts), cut_edges(p_r, n_parts), cut_edges(p_h, n_parts)
        assert hil < 0.75 * rcm < 0.75 * orig, (n_parts, orig, rcm, hil)
    plans = [lib.Plan(p_h, 8, r) for r in range(8)]
    assert sum(pl.n_cut for pl in plans) == 2 * cut_edges(p_h, 8)
    print("%s: cut edges at 8 partitions  original %d  rcm %d  hilbert %d" % (name, cut_edges(p, 8), cut_edges(p_r, 8), cut_edges(p_h, 8)))


def test_hilbert_order_rejects_bad_requests():
    import ctypes as C
    from gocfd_b200 import lib
    l = lib.load()
    ev = np.array([0, 1, 7], dtype=np.int32)
    v = np.zeros(3)
    o = np.zeros(1, dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    assert l.dfr2d_hilbert_order(1, 3, ev.ctypes.data_as(ip), v.ctypes.data_as(dp), v.ctypes.data_as(dp), o.ctypes.data_as(ip)) == 1
    assert b"out of range" in l.dfr2d_last_error(None)
    assert l.dfr2d_hilbert_order(0, 3, None, None, None, None) == 1
