"""Worker of test_two_rank_gloo_dissipation_exchange: the shared-vertex max merge of the PerssonC0 path across two
processes (gloo, CPU only) -- plan lists from the C ABI, the same send/recv pairing bench.py issues over NCCL."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from gocfd_b200 import lib
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D

MESH = sys.argv[3]


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=2, FinalTime=0.2,
                           MaxIterations=10, Gamma=1.4, Minf=0.0, Limiter="persson c0", Kappa=5.0)
    c = Euler(ip, MESH)
    p = c.problem
    assert p.Dissipation
    pl = lib.Plan(p, world, rank)
    rng = np.random.default_rng(11)                    # same stream on every rank: a global per-element scalar
    sigma_k = rng.random(p.K)
    eps_k = rng.random(p.K)
    own = slice(pl.k0, pl.k1)
    sig_v, eps_v = np.zeros(p.NV), np.zeros(p.NV)
    np.maximum.at(sig_v, p.EToV[own].reshape(-1), np.repeat(sigma_k[own], 3))
    np.maximum.at(eps_v, p.EToV[own].reshape(-1), np.repeat(eps_k[own], 3))
    # message per peer: (sigma, eps) interleaved per listed vertex -- the layout of k_vertex_pack
    send = np.empty(2 * len(pl.vertex_ids))
    send[0::2], send[1::2] = sig_v[pl.vertex_ids], eps_v[pl.vertex_ids]
    send_t = torch.from_numpy(send)
    recv_t = torch.empty_like(send_t)
    counts = pl.vertex_counts.tolist()
    ins, outs = list(send_t.split(counts)), list(recv_t.split(counts))
    reqs = []
    for peer in range(world):
        if peer != rank and counts[peer]:
            reqs.append(dist.isend(ins[peer].contiguous(), peer))
            reqs.append(dist.irecv(outs[peer], peer))
    for r in reqs:
        r.wait()
    recv = torch.cat(outs).numpy()
    np.maximum.at(sig_v, pl.vertex_ids, recv[0::2])     # k_vertex_unpack_max
    np.maximum.at(eps_v, pl.vertex_ids, recv[1::2])
    want_s, want_e = np.zeros(p.NV), np.zeros(p.NV)
    np.maximum.at(want_s, p.EToV.reshape(-1), np.repeat(sigma_k, 3))
    np.maximum.at(want_e, p.EToV.reshape(-1), np.repeat(eps_k, 3))
    mine = np.unique(p.EToV[own])
    assert np.array_equal(sig_v[mine], want_s[mine]) and np.array_equal(eps_v[mine], want_e[mine])
    # MergeVertexScalarToElement of the own elements now equals the single-process result
    assert np.array_equal(sig_v[p.EToV[own]].sum(axis=1), want_s[p.EToV[own]].sum(axis=1))
    assert len(pl.vertex_ids) > 0 and int(pl.send_counts.sum()) == pl.n_cut * (4 * p.NpEdge + 3)
    dist.destroy_process_group()
    print("rank %d OK: %d shared vertices, %d cut edges" % (rank, len(pl.vertex_ids), pl.n_cut))


if __name__ == "__main__":
    main()
