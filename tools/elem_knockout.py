"""Timing experiment for kernel 5 (k_elem_ws): the same element kernel built with parts of the consumer work knocked out
(-DDFR2D_WS_KNOCKOUT=bits, see csrc/dfr2d_elem_ws.cuh) to see which part sits on the critical path.  Results of the
knock-out builds are wrong by construction; only the CUDA-event time of the stage_update launches is read.

    for k in 1 2 8 16; do nvcc ... -DDFR2D_WS_KNOCKOUT=$k -o gocfd_b200/csrc/ko/libdfr2d_ko$k.so gocfd_b200/csrc/dfr2d.cu; done
    python tools/elem_knockout.py [--nx 1000 --order 4] [--libs path,path,...]
"""
import argparse
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--libs", default="")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    from gocfd_b200 import lib
    base = lib.LIB_PATH
    libs = [p for p in args.libs.split(",") if p] or [base] + sorted(glob.glob(os.path.join(ROOT, "gocfd_b200", "csrc", "ko", "*.so")))
    c = bench.build_case(args.nx, args.nx, args.order)
    p = c.problem
    out = {"K": int(p.K), "N": int(p.N)}
    for path in libs:
        lib.LIB_PATH, lib._lib = path, None
        dev = lib.Dfr2d(p)
        dev.set_state(c.Q)
        per_rk = [[] for _ in range(5)]
        edge = []
        for step in range(args.steps + 1):
            for rk in range(5):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                dev.stage_sensor(rk)
                dev.stage_prepare(rk)
                ev[0].record()
                dev.stage_edges_interior(rk)
                dev.stage_edges(rk)
                dev.stage_visc(rk)
                ev[1].record()
                dev.stage_update(rk)
                ev[2].record()
                torch.cuda.synchronize()
                if step > 0:
                    per_rk[rk].append(ev[1].elapsed_time(ev[2]))
                    edge.append(ev[0].elapsed_time(ev[1]))
        name = os.path.basename(path)
        out[name] = {"elem_ms_mean": float(np.mean([np.mean(v) for v in per_rk])), "elem_ms_per_rk": [float(np.mean(v)) for v in per_rk],
                     "edge_ms": float(np.mean(edge))}
        print(name, json.dumps(out[name]), flush=True)
        try:
            dev.close()
        except Exception:  # noqa: BLE001  (a knock-out build may have raised the NaN flag)
            pass
    lib.LIB_PATH, lib._lib = base, None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
