#!/usr/bin/env python
"""Benchmark of the 2D Euler DFR stage on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                      CPU arm: the oracle port on the host cores

A "step" is one SSP-RK(5,4) time step = 5 passes of the hot path (stages) over the whole mesh.
Metric: DOF-stage-updates/s = 4 * NpInt * K * 5 * steps / elapsed  (BASELINE.md section 2).
Default workload: config C5 of BASELINE.json -- synthetic 2000x2000x2 = 8M-triangle isentropic
vortex, N=4, Roe flux, global dt, analytic IVortex/Riemann boundaries -- element-partitioned over
the N GPUs with the reference's PartitionMap ranges (strong scaling: the mesh is fixed).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, N)
    "c5": (2000, 2000, 4),      # 8,000,000 triangles, P=4 (north-star target config)
    "c2": (316, 316, 2),        # 199,712 triangles, P=2
    "c3": (2000, 500, 4),       # 2,000,000 triangles, P=4: Sod tube (In/Out/Wall), PerssonC0 sensor + dissipation path
}
DISSIPATION_WORKLOADS = ("c3",)


def algorithmic_bytes(n):
    """Bytes per element per stage of a maximally fused inviscid stage, and the
    share of the two kernels (DESIGN.md 'Kernels')."""
    np_int, np_edge = (n + 1) * (n + 2) // 2, n + 2
    total = 8.0 * (15.2 * np_int + 42 * np_edge + 15)
    elem = 8.0 * (15.2 * np_int + 24 * np_edge + 7.5)
    edge = 8.0 * (18 * np_edge + 7.5)
    return total, elem, edge


def algorithmic_bytes_visc(n):
    """B_visc(N) of SURVEY.md 8d: the inviscid bytes plus the sensor / vertex merge / RT gradient / viscous flux /
    limiter traffic of the PerssonC0 path, per element per stage."""
    np_int, np_edge = (n + 1) * (n + 2) // 2, n + 2
    return algorithmic_bytes(n)[0] + 8.0 * (17.6 * np_int + 84 * np_edge + 18)


def build_case(nx, ny, n, max_iter=10 ** 9, dissipation=False):
    from gocfd_b200.host.euler2d import Euler
    from gocfd_b200.host.input_parameters import InputParameters2D
    from gocfd_b200.host.meshgen import structured_tri_mesh
    if dissipation:
        # config C3 scaled up: the shipped sod-aligned meshes are [0,1] x [0,~0.1] tubes with in/out/wall tags
        ip = InputParameters2D(Title="bench", CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n,
                               FinalTime=1.0e9, MaxIterations=max_iter, Gamma=1.4, Minf=0.0, Limiter="persson c0", Kappa=5.0)
        mesh = structured_tri_mesh(nx, ny, 0.0, 1.0, 0.0, float(ny) / float(nx),
                                   side_tags={"left": "in", "right": "out", "top": "wall", "bottom": "wall"})
        c = Euler(ip, mesh)
        # a raw jump inside a P4 element undershoots to negative density at the edge points (the reference would
        # NaN-panic as well), so the front is smeared over half an element width and kept off the grid lines
        x, _ = c.DFR.solution_xy()
        h = 1.0 / nx
        w = 0.5 * (1.0 - np.tanh((x - (0.5 + 0.3 * h)) / (0.5 * h)))
        for v in range(4):
            c.Q[v] = c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w
        return c
    ip = InputParameters2D(Title="bench", CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=n,
                           FinalTime=1.0e9, MaxIterations=max_iter, Gamma=1.4, Minf=0.1)
    return Euler(ip, structured_tri_mesh(nx, ny, tag="wall"))


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class _DevArray:
    """Expose a raw device pointer to torch through the CUDA array interface."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


DISS_SAMPLE_NX = 600     # 600x150x2 = 180,000 triangles for the (slower, partly serial) PerssonC0 path of the port
CPU_SAMPLE_NX = 500      # 500x500x2 = 500,000 triangles: state >> CPU caches, one RK step ~1 s on 8 cores at N=4


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def cpu_run(n, steps=None, warmup=1, seconds_target=12.0, nx=CPU_SAMPLE_NX, dissipation=False):
    """The CPU arm: oracle/c (C + OpenMP restatement of the reference's stage, phase structure and
    materialised arrays of the Go solver, all host threads) on a bounded sample of the same workload:
    same vortex set-up, order, flux and dt mode on an nx x nx x 2 mesh.  Runs `steps` RK steps, or as
    many as fit in seconds_target when steps is None.  The Go toolchain is absent, so this is a port."""
    from oracle.c_oracle import COracleSolver, threads
    ny = max(1, nx // 4) if dissipation else nx          # the Sod tube keeps its 4:1 aspect
    c = build_case(nx, ny, n, dissipation=dissipation)
    o = COracleSolver(c.problem)
    o.set_state(c.Q)
    o.step(max(1, warmup))
    t0 = time.perf_counter()
    done = 0
    while (steps is not None and done < steps) or (steps is None and (done < 2 or time.perf_counter() - t0 < seconds_target)):
        o.step(1)
        done += 1
    el = time.perf_counter() - t0
    o.close()
    dof = 4 * c.problem.NpInt * c.problem.K * 5 * done
    return {"value": dof / el, "unit": "DOF-stage-updates/s", "cores": threads(), "kind": "port",
            "us_per_element_iteration": el * 1e6 / done / c.problem.K, "ms_per_step": el * 1e3 / done,
            "sample": "C/OpenMP restatement of the Go stage (oracle/c%s), %d threads on %s; %dx%dx2=%d triangles, N=%d, "
                      "%d RK steps in %.1f s; not the Go solver (no Go toolchain)"
                      % (", PerssonC0 path" if dissipation else "", threads(), _cpu_model(), nx, ny, c.problem.K, n, done, el)}


def cpu_baseline(n, dissipation=False):
    return cpu_run(n, nx=DISS_SAMPLE_NX if dissipation else CPU_SAMPLE_NX, dissipation=dissipation)


def run_reference(args, rank):
    if rank != 0:
        return
    nx, ny, n = WORKLOADS[args.workload]
    if args.order >= 0:
        n = args.order
    t_all = time.perf_counter()
    diss = args.workload in DISSIPATION_WORKLOADS
    sample_nx = min(DISS_SAMPLE_NX if diss else CPU_SAMPLE_NX, nx)
    base = cpu_run(n, steps=args.steps, warmup=args.warmup, nx=sample_nx, dissipation=diss)
    line = {
        "impl": "reference", "metric": "DOF-stage-updates/s", "value": base["value"], "unit": "DOF-stage-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s, %dx%dx2 triangles, N=%d, Roe, global dt (the CPU arm steps a bounded "
                               "%dx%dx2 sample of it; throughput is size-independent once the state exceeds the caches)"
                               % (args.workload, "Sod shock tube with PerssonC0 dissipation" if diss else "isentropic vortex",
                                  nx, ny, n, sample_nx, max(1, sample_nx // 4) if diss else sample_nx)},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "DOF-stage-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=0, help="override the mesh size (debug)")
    ap.add_argument("--order", type=int, default=-1, help="override the polynomial order (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from gocfd_b200 import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nx, ny, n = WORKLOADS[args.workload]
    if args.nx:
        nx = ny = args.nx
    if args.order >= 0:
        n = args.order
    t_setup = time.perf_counter()
    diss = args.workload in DISSIPATION_WORKLOADS
    c = build_case(nx, ny, n, dissipation=diss)
    p = c.problem
    assert bool(p.Dissipation) == diss
    dev = lib.Dfr2d(p, n_parts=world, part=rank, device=local_rank)
    stream = torch.cuda.current_stream()
    dev.set_stream(stream.cuda_stream)
    k0, k1 = dev.partition_range()
    setup_s = time.perf_counter() - t_setup

    # state lives in pinned host memory (the Go side would pin c.Q the same way)
    q_host_t = torch.empty((4, p.NpInt, p.K), dtype=torch.float64, pin_memory=True)
    q_host = q_host_t.numpy()
    q_host[...] = c.Q
    dev.set_state(q_host)

    if world > 1:
        # one NCCL all_to_all per exchange point of the stage (1 inviscid, 3 with dissipation) + the MAX allreduce
        def exchange(which):
            sc, rc = dev.exchange_counts(which)
            sp, rp = dev.exchange_buffers(which)
            st = torch.as_tensor(_DevArray(sp, max(sum(sc), 1)), device="cuda")[:sum(sc)]
            rt = torch.as_tensor(_DevArray(rp, max(sum(rc), 1)), device="cuda")[:sum(rc)]
            return lambda async_op=False: dist.all_to_all_single(rt, st, rc, sc, async_op=async_op)
        x_edge = exchange(dev.XCHG_EDGE)
        x_vtx = exchange(dev.XCHG_VERTEX) if diss else None
        x_diss = exchange(dev.XCHG_DISS) if diss else None

        def one_step():
            for rk in range(5):
                if diss:
                    dev.stage_sensor(rk)
                    x_vtx()
                dev.stage_prepare(rk)
                work = x_edge(async_op=True)         # NCCL stream; waits for the pack kernel, not for what follows
                dev.stage_edges_interior(rk)         # interior-edge fluxes overlap the halo transfer
                work.wait()                          # the compute stream waits for the halo (no host block)
                dev.stage_edges(rk)                  # unpack + boundary and cut edges
                if diss:
                    x_diss()
                    dev.stage_visc(rk)
                wp = dev.wavespeed_buffer()
                w = torch.as_tensor(_DevArray(wp, 2), device="cuda")
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
                dev.stage_update(rk)

        def run_steps(k):
            for _ in range(k):
                one_step()
    else:
        def run_steps(k):
            dev.step(k, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # started before the warm-up so that samples exist even for sub-second timed regions
    run_steps(args.warmup)
    barrier()
    l0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    run_steps(args.steps)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = dev.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    dof_per_step = 4 * p.NpInt * p.K * 5
    value = dof_per_step * args.steps / (ms * 1e-3)

    # ---- per-kernel timing of the dominant kernel (k_elem) with CUDA events on the launch stream
    b_total, b_elem, b_edge = algorithmic_bytes(n)
    k_local = k1 - k0
    t_elem, t_edge = [], []
    if world == 1 and not diss:
        evs = []
        for _ in range(2):
            for rk in range(5):
                a, b, cc = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                dev.stage_prepare(rk)
                a.record(stream)
                dev.stage_edges(rk)
                b.record(stream)
                dev.stage_update(rk)
                cc.record(stream)
                evs.append((a, b, cc))
        torch.cuda.synchronize()
        t_edge = [a.elapsed_time(b) for a, b, _ in evs]
        t_elem = [b.elapsed_time(cc) for _, b, cc in evs]
    phases = None
    if world == 1 and diss:
        # the five launches groups of a PerssonC0 stage through the stage API (single partition: nothing to exchange)
        names = ["k_sensor", "k_diss_prepare", "k_edge + RT gradient", "k_visc_edge", "k_elem<N,true>"]
        evs = []
        for _ in range(2):
            for rk in range(5):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                ev[0].record(stream)
                dev.stage_sensor(rk)
                ev[1].record(stream)
                dev.stage_prepare(rk)
                ev[2].record(stream)
                dev.stage_edges(rk)
                ev[3].record(stream)
                dev.stage_visc(rk)
                ev[4].record(stream)
                dev.stage_update(rk)
                ev[5].record(stream)
                evs.append(ev)
        torch.cuda.synchronize()
        phases = {nm: statistics.mean(e[i].elapsed_time(e[i + 1]) for e in evs) for i, nm in enumerate(names)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline = None
    if diss:
        bv = algorithmic_bytes_visc(n)
        ach = bv * p.K * 5 * args.steps / (ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole PerssonC0 stage (k_sensor, k_diss_prepare, k_edge, k_grad_pipe [DMMA], k_visc_edge, k_elem<N,true>)",
                    "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / peak / world, "traffic": None,
                    "peak_source": peak_src, "bytes_per_element_stage": bv, "phase_ms": phases}
    if t_elem:
        te = statistics.mean(t_elem) * 1e-3
        ach = b_elem * k_local / te / 1e9
        roofline = {"bound": "hbm", "kernel": ("k_elem_tma<%d,8> (warp-specialised async-copy pipeline, FP64 DMMA contractions)" if n >= 2 else "k_elem<%d,false>") % n, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": b_elem * k_local, "avg_launch_ms": te * 1e3,
                    "edge_kernel": {"achieved": b_edge * k_local / (statistics.mean(t_edge) * 1e-3) / 1e9,
                                    "avg_launch_ms": statistics.mean(t_edge)},
                    "whole_stage": {"bytes_per_element": b_total,
                                    "achieved": b_total * p.K * 5 * args.steps / (ms * 1e-3) / 1e9,
                                    "frac": b_total * p.K * 5 * args.steps / (ms * 1e-3) / 1e9 / peak / world}}
        prof = os.path.join(ROOT, "profiles", "r01j_traffic.json")
        if os.path.exists(prof):
            try:
                roofline["traffic"] = json.load(open(prof)).get("k_elem_N%d" % n)
            except Exception:
                pass

    # ---- end to end through the C ABI with host buffers: upload state, K steps each returning
    # step info to the host (as Solve does for its progress line), download state.
    e2e = None
    if world == 1:
        q_host[...] = c.Q
        k_e2e = args.steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.set_state(q_host)
        for _ in range(k_e2e):
            info = dev.step(1, sync=True)
        dev.get_state(q_host)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        qb = q_host.nbytes
        e2e = {"value": dof_per_step * k_e2e / el, "unit": "DOF-stage-updates/s",
               "h2d_bytes_per_step": qb / k_e2e, "d2h_bytes_per_step": qb / k_e2e + 40,
               "what": "dfr2d_set_state(pinned host Q) + %d x dfr2d_step(1, info) + dfr2d_get_state" % k_e2e}
    else:
        # every rank uploads its columns, steps with the NCCL exchange, downloads its columns
        q_host[...] = c.Q
        barrier()
        t0 = time.perf_counter()
        dev.set_state(q_host)
        for _ in range(args.steps):
            one_step()
            dev.step_finish(sync=True)
        dev.get_state(q_host)
        barrier()
        el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        qb = 4 * p.NpInt * (k1 - k0) * 8
        e2e = {"value": dof_per_step * args.steps / float(el.item()), "unit": "DOF-stage-updates/s",
               "h2d_bytes_per_step": qb / args.steps, "d2h_bytes_per_step": qb / args.steps + 40,
               "what": "per rank: set_state(own columns) + steps with NCCL halo exchange + step info + get_state"}

    if rank == 0:
        line = {
            "metric": "DOF-stage-updates/s", "value": value, "unit": "DOF-stage-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("%s: Sod shock tube, %dx%dx2=%d triangles, N=%d (NpInt=%d), Roe flux, global dt, PerssonC0 "
                                    "sensor + artificial dissipation, In/Out/Wall boundaries" if diss else
                                    "%s: isentropic vortex, %dx%dx2=%d triangles, N=%d (NpInt=%d), Roe flux, global dt, "
                                    "IVortex+Riemann boundaries") % (args.workload, nx, ny, p.K, n, p.NpInt),
                       "partition": "PartitionMap.Split1D element ranges over %d GPU(s)" % world,
                       "l2": "no flush needed: per-GPU working set %.1f GB >> 126 MB L2"
                             % ((5 * 4 * p.NpInt + 12 * p.NpEdge + 6 * p.NpEdge) * 8 * k_local / 1e9),
                       "setup_s": setup_s},
            "element_stages_per_s": p.K * 5 * args.steps / (ms * 1e-3),
            "us_per_element_iteration": ms * 1e3 / args.steps / p.K,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(n, dissipation=diss)
        print(json.dumps(line))
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
