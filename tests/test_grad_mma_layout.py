"""CPU check of the tensor-core gradient kernel's data layout (k_grad_mma, gocfd_b200/csrc/dfr2d_grad_mma.cuh).

The operator table comes from the library's own host-side builder (dfr2d_grad_mma_table); the DMMA.8x8x4 lane
semantics (A: lane l holds A[l/4][l%4]; B: lane l holds B[k=l%4][n=l/4]; C: lane l holds C[l/4][2(l%4)], [..+1]) and the
kernel's shared-memory indexing are replayed in numpy, element tile by element tile, and compared with the dense
statement of GetSolutionGradientUsingRTElement (euler.go:864-918): Grad = Div . (Metric (.) U), Diss = Epsilon (.) Grad.
No GPU involved: this pins the block decomposition (two interior blocks + three edge blocks, zero padding per block),
the produced-row map and the Bary rows, which are what a wrong index would break.
"""
import numpy as np
import pytest

from gocfd_b200 import lib
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh

E, SE = 32, 36


def _dims(n):
    ni, ned = (n + 1) * (n + 2) // 2, n + 2
    nf = (n + 2) * (n + 4)
    nout = ni + 3 * ned
    ntail = 1 if nout % 8 == 1 else 0                    # a single ragged row: plain DFMA (grad_tail), not a padded m-tile
    mt, ki, ke = (nout - ntail + 7) // 8, (ni + 3) // 4, (ned + 3) // 4
    return ni, ned, nf, nout, mt, ki, ke, 2 * ki + 3 * ke, 4 * ki + 12 * ke


def _dmma(acc, a_lane, b_lane):
    """acc[32][2] += A . B with the m8n8k4 fragment layouts."""
    a = np.zeros((8, 4))
    b = np.zeros((4, 8))
    for l in range(32):
        a[l // 4, l % 4] = a_lane[l]
        b[l % 4, l // 4] = b_lane[l]
    c = a @ b
    for l in range(32):
        acc[l, 0] += c[l // 4, 2 * (l % 4)]
        acc[l, 1] += c[l // 4, 2 * (l % 4) + 1]


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_grad_mma_tile_replay_matches_dense_gradient(n):
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=n, FinalTime=1.0,
                           MaxIterations=10, Gamma=1.4, Minf=0.5, Limiter="PerssonC0", Kappa=5.0)
    c = Euler(ip, structured_tri_mesh(2, 2, tag="far"))
    p = c.problem
    ni, ned, nf, nout, mt_n, ki, ke, ks_n, urows = _dims(n)
    ntail = 1 if nout % 8 == 1 else 0
    assert (ni, ned, nf) == (p.NpInt, p.NpEdge, p.NpFlux)
    table = lib.grad_mma_table(p)
    nfrag = mt_n * ks_n * 32
    assert table.size == nfrag + 8 * mt_n * 3
    div, bary = np.asarray(p.Div), np.asarray(p.Bary)

    rng = np.random.default_rng(n)
    # one tile of 32 elements, one conserved variable: distinct values, metrics and vertex epsilons
    u_int = rng.standard_normal((ni, E))
    u_edge = rng.standard_normal((3, ned, E))
    met = rng.standard_normal((5, 2, E))               # block r, x|y
    ev = rng.random((3, E))

    # dense statement
    un = np.concatenate([u_int, u_int, u_edge.reshape(3 * ned, E)])
    mxr = np.concatenate([np.repeat(met[0:1, 0], ni, 0), np.repeat(met[1:2, 0], ni, 0)] +
                         [np.repeat(met[2 + e:3 + e, 0], ned, 0) for e in range(3)])
    myr = np.concatenate([np.repeat(met[0:1, 1], ni, 0), np.repeat(met[1:2, 1], ni, 0)] +
                         [np.repeat(met[2 + e:3 + e, 1], ned, 0) for e in range(3)])
    eps = bary @ ev
    want_x = (div @ (mxr * un)) * eps
    want_y = (div @ (myr * un)) * eps

    # kernel replay: shared-memory image of U
    s_u = np.full((urows, SE), np.nan)
    for i in range(4 * ki):
        s_u[i, :E] = u_int[i] if i < ni else 0.0
    for le in range(3):
        for i in range(4 * ke):
            s_u[4 * ki + le * 4 * ke + i, :E] = u_edge[le, i] if i < ned else 0.0
    blk_ks = [ki, ki, ke, ke, ke]
    blk_k0 = [0, ki, 2 * ki, 2 * ki + ke, 2 * ki + 2 * ke]
    blk_u0 = [0, 0, 4 * ki, 4 * ki + 4 * ke, 4 * ki + 8 * ke]
    got_x = np.zeros((nf, E))
    got_y = np.zeros((nf, E))
    lanes = np.arange(32)
    fr, fc = lanes >> 2, lanes & 3
    for mt in range(mt_n):
        for nt in range(4):
            gx = np.zeros((32, 2))
            gy = np.zeros((32, 2))
            e0 = 8 * nt + 2 * fc
            for r in range(5):
                s = np.zeros((32, 2))
                for ks in range(blk_ks[r]):
                    b_lane = s_u[blk_u0[r] + 4 * ks + fc, 8 * nt + fr]
                    a_lane = table[(mt * ks_n + blk_k0[r] + ks) * 32 + lanes]
                    _dmma(s, a_lane, b_lane)
                for cc in range(2):
                    gx[:, cc] += met[r, 0, e0 + cc] * s[:, cc]
                    gy[:, cc] += met[r, 1, e0 + cc] * s[:, cc]
            for l in range(32):
                m = 8 * mt + fr[l]
                if m >= nout - ntail:
                    continue
                row = m if m < ni else m + ni
                b3 = table[nfrag + 3 * m:nfrag + 3 * m + 3]
                for cc in range(2):
                    e = e0[l] + cc
                    epsv = b3[0] * ev[0, e] + b3[1] * ev[1, e] + b3[2] * ev[2, e]
                    got_x[row, e] = gx[l, cc] * epsv
                    got_y[row, e] = gy[l, cc] * epsv
    # the ragged rows: grad_tail's block-wise DFMA sums over the same shared-memory image (part 0: blocks 0, 2, 3;
    # part 1: blocks 1, 4; one shuffle adds the halves), Div / Bary taken from the dense operators
    blk_cols = [ni, ni, ned, ned, ned]
    blk_col0 = [0, ni, 2 * ni, 2 * ni + ned, 2 * ni + 2 * ned]
    for t in range(ntail):
        m = 8 * mt_n + t
        row = m + ni
        for e in range(E):
            parts = []
            for blocks in ((0, 2, 3), (1, 4)):
                gxp = gyp = 0.0
                for r in blocks:
                    sr = sum(div[row, blk_col0[r] + j] * s_u[blk_u0[r] + j, e] for j in range(blk_cols[r]))
                    gxp += met[r, 0, e] * sr
                    gyp += met[r, 1, e] * sr
                parts.append((gxp, gyp))
            epsv = bary[row] @ ev[:, e]
            got_x[row, e] = (parts[0][0] + parts[1][0]) * epsv
            got_y[row, e] = (parts[0][1] + parts[1][1]) * epsv
    rows = list(range(ni)) + list(range(2 * ni, nf))
    scale = np.abs(want_x[rows]).max()
    assert np.abs(got_x[rows] - want_x[rows]).max() < 1e-12 * scale
    assert np.abs(got_y[rows] - want_y[rows]).max() < 1e-12 * np.abs(want_y[rows]).max()
    # rows the kernel never writes are the duplicate interior block only
    assert not got_x[ni:2 * ni].any()


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_elem_mma_diss_tile_replay(n):
    """k_elem_mma_diss (gocfd_b200/csrc/dfr2d_elem_mma_diss.cuh): RHS = -(1/J) DivInt . F, then the limiter
    Vinv -> mode scaling -> V, replayed with the library's own fragment table and the kernel's smem row indexing
    (rows [0, QROWS) of the F block are reused for RHS and for the scaled modes), against the dense statement of
    RHSInternalPoints + limitAndFilterSolution (euler.go:665-699, dissipation.go:606-622)."""
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="Freestream", PolynomialOrder=n, FinalTime=1.0,
                           MaxIterations=10, Gamma=1.4, Minf=0.5, Limiter="PerssonC0", Kappa=5.0)
    c = Euler(ip, structured_tri_mesh(2, 2, tag="far"))
    p = c.problem
    ni, nf = p.NpInt, p.NpFlux
    m1, k1, m3, k3 = (ni + 7) // 8, (nf + 3) // 4, (ni + 7) // 8, (ni + 3) // 4
    qrows, frows = 4 * k3, 4 * k1
    table = lib.mma_diss_table(p)
    assert table.size == (m1 * k1 + 2 * m3 * k3) * 32
    divint, vinv, v, mf = np.asarray(p.DivInt), np.asarray(p.Vinv), np.asarray(p.V), np.asarray(p.ModeFilter)
    rng = np.random.default_rng(10 + n)
    f = rng.standard_normal((nf, E))
    mooj = -1.0 / (0.5 + rng.random(E))
    oma = rng.random(E)                                   # 1 - sin(pi sigma / 2)
    # dense statement
    rhs = (divint @ f) * mooj
    uh = vinv @ rhs
    uh[1:] *= (mf[1:, None] * oma[None, :])
    want = v @ uh
    # replay
    lanes = np.arange(32)
    fr, fc = lanes >> 2, lanes & 3
    s_f = np.zeros((frows, SE))
    s_f[:nf, :E] = f

    def product(frag0, mt_n, ks_n):
        """C[mt][nt] fragments of (operator at table offset frag0) . s_f[rows < 4 ks_n]."""
        acc = [[np.zeros((32, 2)) for _ in range(4)] for _ in range(mt_n)]
        for ks in range(ks_n):
            for nt in range(4):
                b_lane = s_f[4 * ks + fc, 8 * nt + fr]
                for mt in range(mt_n):
                    a_lane = table[(frag0 + mt * ks_n + ks) * 32 + lanes]
                    _dmma(acc[mt][nt], a_lane, b_lane)
        return acc

    def store_rows(acc, mt_n, scale=None):
        for mt in range(mt_n):
            for l in range(32):
                i = 8 * mt + fr[l]
                if i >= qrows:
                    continue
                for nt in range(4):
                    e0 = 8 * nt + 2 * fc[l]
                    for cc in range(2):
                        val = acc[mt][nt][l, cc]
                        if scale is not None:
                            val = scale(i, e0 + cc, val)
                        s_f[i, e0 + cc] = val

    c1 = product(0, m1, k1)
    store_rows(c1, m1, lambda i, e, val: val * mooj[e])
    c3 = product(m1 * k1, m3, k3)
    store_rows(c3, m3, lambda i, e, val: val * ((mf[i] if 1 <= i < ni else 0.0) * oma[e]) if i >= 1 else val)
    c4 = product(m1 * k1 + m3 * k3, m3, k3)
    got = np.zeros((ni, E))
    for mt in range(m3):
        for l in range(32):
            i = 8 * mt + fr[l]
            if i < ni:
                for nt in range(4):
                    e0 = 8 * nt + 2 * fc[l]
                    got[i, e0] = c4[mt][nt][l, 0]
                    got[i, e0 + 1] = c4[mt][nt][l, 1]
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
