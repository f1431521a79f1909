"""A/B of the kernel variants of the PerssonC0 path on one GPU (variant 9 = tensor-core element kernel k_elem_mma_diss;
the rest are the RT-gradient kernels) (DFR2D_GRAD_KERNEL=1: constant-operand DFMA
k_grad, =2: DMMA k_grad_mma, =3: pipelined persistent DMMA k_grad_pipe with DFR2D_GRAD_MG=3|2): whole-step time, the stage_edges phase (k_edge + k_grad) per RK stage, and the
relative L2 difference of the two states after the same number of steps.  One JSON line on stdout.

    python tools/grad_kernel_ab.py [--nx 2000 --ny 500 --order 4 --steps 6]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--variants", default="1,3,9", help="subset of the variant labels to run (1 is the reference state)")
    ap.add_argument("--lib", default="", help="alternative build of libdfr2d.so (compile-time variants)")
    args = ap.parse_args()
    import torch
    import bench
    from gocfd_b200 import lib
    if args.lib:
        lib.LIB_PATH = os.path.abspath(args.lib)
    c = bench.build_case(args.nx, args.ny, args.order, dissipation=True)
    p = c.problem
    out = {"K": int(p.K), "N": int(p.N), "steps": args.steps}
    states = {}
    # label -> (DFR2D_GRAD_KERNEL, DFR2D_GRAD_MG, DFR2D_GRAD_SKEW_NS, DFR2D_DISS_ELEM_KERNEL)
    variants = {1: ("1", "3", "0", "1"), 2: ("2", "3", "0", "1"), 3: ("3", "3", "0", "1"), 4: ("3", "2", "0", "1"),
                5: ("3", "3", "1500", "1"), 6: ("3", "3", "3000", "1"), 7: ("3", "3", "5000", "1"), 8: ("3", "2", "3000", "1"),
                9: ("3", "3", "0", "3"), 10: ("3", "3", "0", "5"), 11: ("3", "2", "0", "5"), 12: ("4", "3", "0", "5"), 13: ("4", "2", "0", "5")}
    variants = {k: v for k, v in variants.items() if str(k) in args.variants.split(",")}
    for gk in variants:
        (os.environ["DFR2D_GRAD_KERNEL"], os.environ["DFR2D_GRAD_MG"], os.environ["DFR2D_GRAD_SKEW_NS"],
         os.environ["DFR2D_DISS_ELEM_KERNEL"]) = variants[gk]
        dev = lib.Dfr2d(p)
        dev.set_state(c.Q)
        dev.step(2, sync=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dev.step(args.steps, sync=False)
        e1.record()
        torch.cuda.synchronize()
        ms_step = e0.elapsed_time(e1) / args.steps
        states[gk] = dev.get_state()
        # phase timing of one more step through the stage API (single partition: no exchange in between)
        ph = []
        for rk in range(5):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            dev.stage_sensor(rk)
            ev[1].record()
            dev.stage_prepare(rk)
            ev[2].record()
            dev.stage_edges(rk)
            ev[3].record()
            dev.stage_visc(rk)
            ev[4].record()
            dev.stage_update(rk)
            ev[5].record()
            ph.append(ev)
        dev.step_finish(sync=True)
        torch.cuda.synchronize()
        names = ["sensor", "prepare", "edges+grad", "visc", "update"]
        out["grad_kernel_%d" % gk] = {
            "ms_per_step": ms_step, "ms_per_stage": ms_step / 5,
            "phase_ms_mean": {nm: float(np.mean([e[i].elapsed_time(e[i + 1]) for e in ph])) for i, nm in enumerate(names)},
            "finite": bool(np.isfinite(states[gk]).all()),
        }
        dev.close()
    out["variants"] = {"grad_kernel_%d" % k: "DFR2D_GRAD_KERNEL=%s DFR2D_GRAD_MG=%s DFR2D_GRAD_SKEW_NS=%s DFR2D_DISS_ELEM_KERNEL=%s" % v for k, v in variants.items()}
    for k in [k for k in variants if k != 1 and 1 in variants]:
        out["rel_l2_state_%d_vs_1" % k] = float(np.linalg.norm(states[1] - states[k]) / np.linalg.norm(states[1]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
