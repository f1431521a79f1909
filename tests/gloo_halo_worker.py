"""Worker of test_two_rank_gloo_exchange: one process per partition, gloo backend, CPU only."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from gocfd_b200 import lib
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ip = InputParameters2D(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=2, FinalTime=1.0,
                           MaxIterations=10, Gamma=1.4, Minf=0.1)
    c = Euler(ip, structured_tri_mesh(16, 9))
    p = c.problem
    pl = lib.Plan(p, world, rank)
    ne, per = p.NpEdge, 4 * p.NpEdge
    qface = np.stack([p.FluxEdgeInterp @ c.Q[n] for n in range(4)])        # global, identical on every rank
    mine = np.zeros((4, 3 * ne, pl.Kp))
    mine[:, :, :pl.K] = qface[:, :, pl.k0:pl.k1]
    send = np.empty(pl.n_cut * per)
    for ci in range(pl.n_cut):
        send[ci * per:(ci + 1) * per] = mine[:, pl.send_row0[ci] + np.arange(ne), pl.send_elem[ci]].reshape(-1)
    send_t = torch.from_numpy(send)
    recv_t = torch.empty(int(pl.recv_counts.sum()), dtype=torch.float64)
    outs = list(recv_t.split(pl.recv_counts.tolist()))
    ins = list(send_t.split(pl.send_counts.tolist()))
    # gloo has no all_to_all: the same exchange as point-to-point sends (NCCL uses all_to_all_single in bench.py)
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if ins[peer].numel():
            reqs.append(dist.isend(ins[peer].contiguous(), peer))
        if outs[peer].numel():
            reqs.append(dist.irecv(outs[peer], peer))
    for r in reqs:
        r.wait()
    recv = torch.cat(outs).numpy() if outs else np.empty(0)
    for ci in range(pl.n_cut):
        mine[:, pl.recv_row0[ci] + np.arange(ne), pl.recv_col[ci]] = recv[ci * per:(ci + 1) * per].reshape(4, ne)
    col2glob = np.concatenate([np.arange(pl.k0, pl.k1), pl.ghost_global])
    ge = pl.global_edge
    sh = np.flatnonzero(p.edge_nconn[ge] == 2)
    i = np.arange(ne)
    rows_r = p.edge_numR[ge][sh][:, None] * ne + (ne - 1 - i)[None, :]
    rows_l = p.edge_numL[ge][:, None] * ne + i[None, :]
    for n in range(4):
        assert np.array_equal(mine[n][rows_r, pl.kR[sh][:, None]], qface[n][rows_r, col2glob[pl.kR[sh]][:, None]])
        assert np.array_equal(mine[n][rows_l, pl.kL[:, None]], qface[n][rows_l, col2glob[pl.kL][:, None]])
    # global dt: MAX-allreduce of the per-partition wave-speed maxima equals the global maximum
    local_max = torch.tensor([float(rank + 1) * 3.25, 0.0], dtype=torch.float64)
    dist.all_reduce(local_max, op=dist.ReduceOp.MAX)
    assert local_max[0].item() == world * 3.25
    assert pl.n_cut > 0 and pl.G > 0
    dist.destroy_process_group()
    print("rank %d OK: %d cut edges, %d ghost columns" % (rank, pl.n_cut, pl.G))


if __name__ == "__main__":
    main()
