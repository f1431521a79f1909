"""Worker of test_peer_connected_processes_match_single_partition: one process per partition; mailbox descriptions
travel over gloo (CPU), the data path is CUDA IPC peer memory.  argv: rank world diss port out_prefix"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build(diss):
    from gocfd_b200.host.euler2d import Euler
    from gocfd_b200.host.input_parameters import InputParameters2D
    from gocfd_b200.host.meshgen import structured_tri_mesh
    base = dict(CFL=1.0, FluxType="Roe", Minf=0.8, Gamma=1.4, Alpha=1.25, FinalTime=50.0, MaxIterations=1000)
    if diss:
        base.update(PolynomialOrder=2, InitType="shocktube", CFL=2.0, FinalTime=0.2, Limiter="persson c0", Kappa=5.0)
        c = Euler(InputParameters2D(**base), os.path.join(ROOT, "tests", "golden", "meshes", "sod-aligned-100pts.su2"))
        x, _ = c.DFR.solution_xy()
        w = 0.5 * (1.0 - np.tanh((x - 0.503) / 0.004))
        c.Q = np.stack([c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w for v in range(4)])
        return c
    base.update(PolynomialOrder=3, InitType="IVortex")
    return Euler(InputParameters2D(**base), structured_tri_mesh(16, 12))


def main():
    rank, world, diss, port, out = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3])), sys.argv[4], sys.argv[5]
    import torch
    import torch.distributed as dist
    from gocfd_b200 import lib
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=world)
    c = build(diss)
    dev = lib.Dfr2d(c.problem, n_parts=world, part=rank, device=rank % torch.cuda.device_count())
    blobs = [None] * world
    dist.all_gather_object(blobs, dev.peer_export())
    dev.peer_connect(blobs)
    dist.barrier()                      # nobody puts before every mailbox is mapped
    dev.set_state(c.Q)
    dev.step(1)
    info = dev.step(3)
    k0, k1 = dev.partition_range()
    np.savez("%s_%d.npz" % (out, rank), q=dev.get_state()[:, :, k0:k1], k0=k0, k1=k1, time=info["time"], steps=info["steps"])
    dist.barrier()                      # keep the mailbox alive until every partner has finished its last stage
    dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
