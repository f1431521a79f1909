#!/usr/bin/env python
"""Benchmark of the 2D Euler DFR stage on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                      CPU arm: the oracle port on the host cores

A "step" is one SSP-RK(5,4) time step = 5 passes of the hot path (stages) over the whole mesh.
Metric: DOF-stage-updates/s = 4 * NpInt * K * 5 * steps / elapsed  (BASELINE.md section 2).
Default workload: config C5 of BASELINE.json -- synthetic 2000x2000x2 = 8M-triangle isentropic
vortex, N=4, Roe flux, global dt, analytic IVortex/Riemann boundaries -- element-partitioned over
the N GPUs with the reference's PartitionMap ranges (strong scaling: the mesh is fixed).

Multi-GPU drivers (N > 1), all over the same kernels and the same partition plan:
  peer        (default) one process per GPU; the partitions exchange halo and wave-speed maxima by themselves over
              CUDA-IPC peer memory (csrc/dfr2d_peer.cuh): every rank just calls dfr2d_step(), no NCCL and no Python in
              the stage loop.  torch.distributed only carries the mailbox descriptions at start-up and the timing
              reductions.
  nccl        round 1's host-driven protocol: Python calls the stage API and moves the bytes with NCCL
              (all_to_all_single + all_reduce(MAX)); kept as the comparison line.
  multi_step  ONE process (rank 0) drives all N GPUs through dfr2d_multi_step -- the shape the Go controller goroutine
              needs.  Reported under "multi_step" of the same JSON line, or alone with
              `python bench.py --gpus N --driver multi_step` (no torchrun).
The default line also carries the P=2 config (C2) and the PerssonC0 shock-capturing config (C3) under "also".
"""
import argparse
import json
import os
import shutil
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


def build_case(nx, ny, n, max_iter=10 ** 9, dissipation=False, window=None):
    """The host-side problem of a workload.  window = (n_parts, part): only the partition's window of the mesh is built
    (gocfd_b200/host/window.py) and the result carries c.window = (K_global, k_offset)."""
    from gocfd_b200.host.euler2d import Euler
    from gocfd_b200.host.input_parameters import InputParameters2D
    from gocfd_b200.host.meshgen import structured_tri_mesh
    from gocfd_b200.host.window import structured_window_case
    if dissipation:
        # config C3 scaled up: the shipped sod-aligned meshes are [0,1] x [0,~0.1] tubes with in/out/wall tags
        ip = InputParameters2D(Title="bench", CFL=1.0, FluxType="Roe", InitType="shocktube", PolynomialOrder=n,
                               FinalTime=1.0e9, MaxIterations=max_iter, Gamma=1.4, Minf=0.0, Limiter="persson c0", Kappa=5.0)
        box = (0.0, 1.0, 0.0, float(ny) / float(nx))
        mesh_kw = dict(side_tags={"left": "in", "right": "out", "top": "wall", "bottom": "wall"})
    else:
        ip = InputParameters2D(Title="bench", CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=n,
                               FinalTime=1.0e9, MaxIterations=max_iter, Gamma=1.4, Minf=0.1)
        box = (-10.0, 10.0, -10.0, 10.0)
        mesh_kw = dict(tag="wall")
    if window is None:
        c = Euler(ip, structured_tri_mesh(nx, ny, *box, **mesh_kw))
        c.window = None
    else:
        c, c.window = structured_window_case(lambda m: Euler(ip, m), nx, ny, window[0], window[1], *box, **mesh_kw)
    if dissipation:
        # a raw jump inside a P4 element undershoots to negative density at the edge points (the reference would
        # NaN-panic as well), so the front is smeared over half an element width and kept off the grid lines
        x, _ = c.DFR.solution_xy()
        h = 1.0 / nx
        w = 0.5 * (1.0 - np.tanh((x - (0.5 + 0.3 * h)) / (0.5 * h)))
        for v in range(4):
            c.Q[v] = c.FSOut.Qinf[v] + (c.FSIn.Qinf[v] - c.FSOut.Qinf[v]) * w
    return c


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


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def go_probe():
    """SURVEY.md 7.3(1) / BASELINE.md section 3: time the real Go solver if the box can build it."""
    go = shutil.which("go")
    if go is None:
        return {"go": None, "status": "Go reference not runnable on this box: no `go` on PATH (and gonum / avs are not vendored); "
                                      "the CPU arm is the C/OpenMP restatement (kind: port)"}
    try:
        ver = subprocess.run([go, "version"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        ver = "go version failed: %r" % (e,)
    return {"go": go, "version": ver,
            "status": "Go toolchain present but the reference's module dependencies (gonum v0.16.0, gonum/netlib, notargets/avs) "
                      "are neither vendored nor fetchable (no network): reference not built; CPU arm is the port"}


def cpu_run(n, steps=None, warmup=1, seconds_target=12.0, nx=CPU_SAMPLE_NX, ny=None, dissipation=False, threads_wanted=None):
    """The CPU arm: oracle/c (C + OpenMP restatement of the reference's stage, phase structure and
    materialised arrays of the Go solver) on all host cores: same set-up, order, flux and dt mode on an nx x ny x 2 mesh.
    Runs `steps` RK steps, or as many as fit in seconds_target when steps is None.  The thread count is set explicitly
    (torchrun exports OMP_NUM_THREADS=1).  The Go toolchain is absent, so this is a port."""
    from oracle.c_oracle import COracleSolver, set_threads
    nthreads = set_threads(threads_wanted or host_cores())
    if ny is None:
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
    rho = o.get_state()[0]
    o.close()
    dof = 4 * c.problem.NpInt * c.problem.K * 5 * done
    return {"value": dof / el, "unit": "DOF-stage-updates/s", "cores": nthreads, "kind": "port",
            "us_per_element_iteration": el * 1e6 / done / c.problem.K, "ms_per_step": el * 1e3 / done,
            "triangles": int(c.problem.K), "steps_timed": done, "l2_rho": float(np.sqrt((rho * rho).sum())),
            "sample": "C/OpenMP restatement of the Go stage (oracle/c%s), %d threads on %s; %dx%dx2=%d triangles, N=%d, "
                      "%d RK steps in %.1f s; not the Go solver (no Go toolchain)"
                      % (", PerssonC0 path" if dissipation else "", nthreads, _cpu_model(), nx, ny, c.problem.K, n, done, el)}


def cpu_baseline(n, dissipation=False):
    return cpu_run(n, nx=DISS_SAMPLE_NX if dissipation else CPU_SAMPLE_NX, dissipation=dissipation)


def run_reference(args, rank):
    """--impl reference: the CPU arm on all host threads.  When the host has the memory, the timed steps run on the SAME mesh
    as the GPU arm (same_config: true); the bounded 500K-triangle sample is timed next to it to show that the rate does not
    depend on the mesh size once the state exceeds the caches."""
    if rank != 0:
        return
    nx, ny, n = WORKLOADS[args.workload]
    if args.order >= 0:
        n = args.order
    if args.nx:
        nx = ny = args.nx
    t_all = time.perf_counter()
    diss = args.workload in DISSIPATION_WORKLOADS
    if args.cpu_full_worker:
        print(json.dumps(cpu_run(n, steps=args.steps, warmup=1, nx=nx, ny=ny, dissipation=diss)))
        return
    np_int, np_flux = (n + 1) * (n + 2) // 2, (n + 2) * (n + 4)
    k_full = 2 * nx * ny
    # materialised arrays of the port (Go layout): 7 registers + F_RT_DOF + Q_Face + edge stores, + the host-side problem
    need_gb = k_full * 8.0 * (4 * np_int * 8 + 4 * np_flux * (3 if diss else 1) + 60 * (n + 2) + 200) / 1e9
    sample_nx = min(DISS_SAMPLE_NX if diss else CPU_SAMPLE_NX, nx)
    sample_ny = max(1, sample_nx // 4) if diss else sample_nx
    sample = cpu_run(n, steps=None, warmup=1, seconds_target=8.0, nx=sample_nx, ny=sample_ny, dissipation=diss)
    full = None
    budget_s = 150.0
    est_step_s = sample["us_per_element_iteration"] * 1e-6 * k_full
    if not args.sample_only and k_full > sample["triangles"] and _mem_available_gb() > 1.5 * need_gb + 8:
        steps_full = int(max(2, min(args.steps, (budget_s - 25.0) / max(est_step_s, 1e-3) - 1)))
        # in a child process: a full-size run that dies (memory) must not take the line with it
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-full-worker",
                                  "--workload", args.workload, "--steps", str(steps_full), "--order", str(n),
                                  "--nx", str(args.nx)], capture_output=True, text=True, timeout=600, cwd=ROOT)
            if out.returncode == 0:
                full = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            full = None
    base = full or sample
    assert base["cores"] > 1 or host_cores() == 1, "the CPU arm must use all host cores"
    line = {
        "impl": "reference", "metric": "DOF-stage-updates/s", "value": base["value"], "unit": "DOF-stage-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s, %dx%dx2 triangles, N=%d, Roe, global dt" % (
                       args.workload, "Sod shock tube with PerssonC0 dissipation" if diss else "isentropic vortex", nx, ny, n),
                   "same_config": full is not None or k_full <= sample["triangles"],
                   "timed_on": "%d triangles, %d RK steps" % (base["triangles"], base["steps_timed"]),
                   "size_independence": {"sample_%d_triangles" % sample["triangles"]: sample["value"],
                                         **({"full_%d_triangles" % full["triangles"]: full["value"]} if full else {})}},
        "cpu_baseline": base,
        "go_reference": go_probe(),
        "e2e": {"value": base["value"], "unit": "DOF-stage-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank's partition: the three drivers, the timing rules of the contract, roofline, e2e, checksum."""

    def __init__(self, args, workload, rank, world, local_rank, gloo):
        import torch
        import torch.distributed as dist
        from gocfd_b200 import lib
        self.torch, self.dist, self.lib = torch, dist, lib
        self.args, self.workload, self.rank, self.world, self.local_rank, self.gloo = args, workload, rank, world, local_rank, gloo
        nx, ny, n = WORKLOADS[workload]
        if args.nx and workload == args.workload:
            nx = ny = args.nx
        if args.order >= 0 and workload == args.workload:
            n = args.order
        self.nx, self.ny, self.n = nx, ny, n
        self.diss = workload in DISSIPATION_WORKLOADS
        t0 = time.perf_counter()
        # one process per GPU: every rank builds only its window of the mesh (no global problem anywhere); a single
        # process (N = 1, or the multi_step driver) owns the whole mesh like the Go host does
        windowed = world > 1 and args.driver != "multi_step" and not args.global_problem
        self.c = build_case(nx, ny, n, dissipation=self.diss, window=(world, rank) if windowed else None)
        self.p = self.c.problem
        assert bool(self.p.Dissipation) == self.diss
        self.build_s = time.perf_counter() - t0
        self.stream = torch.cuda.current_stream()
        self.k_global = self.c.window[0] if self.c.window else self.p.K
        self.k_off = self.c.window[1] if self.c.window else 0
        self.dof_per_step = 4 * self.p.NpInt * self.k_global * 5
        self._global_case = None

    # ---- per-process partition (drivers peer / nccl) ---------------------------------------------------------------
    def open(self):
        torch, dist, lib = self.torch, self.dist, self.lib
        t0 = time.perf_counter()
        self.dev = lib.Dfr2d(self.p, n_parts=self.world, part=self.rank, device=self.local_rank, window=self.c.window)
        self.dev.set_stream(self.stream.cuda_stream)
        self.k0, self.k1 = self.dev.partition_range()
        self.create_s = time.perf_counter() - t0
        # state lives in pinned host memory (the Go side would pin c.Q the same way)
        self.q_host_t = torch.empty((4, self.p.NpInt, self.p.K), dtype=torch.float64, pin_memory=True)
        self.q_host = self.q_host_t.numpy()
        self.q_host[...] = self.c.Q
        self.dev.set_state(self.q_host)
        self.peer_error = None
        if self.world > 1:
            try:
                blobs = [None] * self.world
                dist.all_gather_object(blobs, self.dev.peer_export(), group=self.gloo)
                self.dev.peer_connect(blobs)
            except Exception as e:  # noqa: BLE001  (no CUDA IPC in this container: fall back to the NCCL-moved exchange)
                self.peer_error = repr(e)
            flags = [None] * self.world
            dist.all_gather_object(flags, self.peer_error, group=self.gloo)
            bad = [f for f in flags if f]
            if bad:                      # someone could not map a mailbox: everybody uses the NCCL-moved exchange
                self.peer_error = bad[0]
                self.dev.peer_enable(False)
            dist.barrier(group=self.gloo)        # nobody puts before every mailbox is mapped
            self._nccl_setup()

    def _nccl_setup(self):
        dev, dist, torch = self.dev, self.dist, self.torch

        def exchange(which):
            sc, rc = dev.exchange_counts(which)
            sp, rp = dev.exchange_buffers(which)
            st = torch.as_tensor(_DevArray(sp, max(sum(sc), 1)), device="cuda")[:sum(sc)]
            rt = torch.as_tensor(_DevArray(rp, max(sum(rc), 1)), device="cuda")[:sum(rc)]
            return lambda async_op=False: dist.all_to_all_single(rt, st, rc, sc, async_op=async_op)
        self.x_edge = exchange(dev.XCHG_EDGE)
        self.x_vtx = exchange(dev.XCHG_VERTEX) if self.diss else None
        self.x_diss = exchange(dev.XCHG_DISS) if self.diss else None

    def nccl_step(self):
        """Round 1's host-driven stage protocol: one NCCL all_to_all per exchange point + the MAX allreduce."""
        dev, dist, torch = self.dev, self.dist, self.torch
        for rk in range(5):
            if self.diss:
                dev.stage_sensor(rk)
                self.x_vtx()
            dev.stage_prepare(rk)
            work = self.x_edge(async_op=True)    # NCCL stream; waits for the pack kernel, not for what follows
            dev.stage_edges_interior(rk)         # interior-edge fluxes overlap the halo transfer
            work.wait()                          # the compute stream waits for the halo (no host block)
            dev.stage_edges(rk)                  # unpack + boundary and cut edges
            if self.diss:
                self.x_diss()
                dev.stage_visc(rk)
            w = torch.as_tensor(_DevArray(dev.wavespeed_buffer(), 2), device="cuda")
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            dev.stage_update(rk)

    def run_steps(self, k, driver):
        if self.world == 1 or driver == "peer":
            self.dev.step(k, sync=False)
        else:
            for _ in range(k):
                self.nccl_step()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_list(self, xs, op):
        if self.world == 1:
            return [float(v) for v in xs]
        t = self.torch.tensor([float(v) for v in xs], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return [float(v) for v in t.tolist()]

    def timed(self, driver, steps, warmup):
        """W untimed steps, then EXACTLY `steps` steps between CUDA events on the launch stream, barrier + synchronize on
        both sides, max over ranks."""
        torch = self.torch
        if self.world > 1:
            self.dev.peer_enable(driver == "peer")
        self.run_steps(warmup, driver)
        self.barrier()
        l0 = self.dev.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record(self.stream)
        self.run_steps(steps, driver)
        ev1.record(self.stream)
        self.barrier()
        ms = self.reduce_max(ev0.elapsed_time(ev1))
        launches = self.dev.launch_count() - l0
        if self.world > 1:
            launches = int(self.reduce_list([launches], self.dist.ReduceOp.SUM)[0])
        return ms, launches

    def kernel_times(self, driver):
        """CUDA-event time of the interior-edge kernel and the element kernel of every stage of two more steps, on this
        rank (every rank runs the same stages, so the exchange protocol stays in lock step)."""
        torch, dev = self.torch, self.dev
        if self.world > 1:
            dev.peer_enable(driver == "peer")
        evs = []
        # (PerssonC0 path: the interior edges run inside stage_edges, after the RT gradient, with the viscous flux fused in;
        # stage_visc is the boundary / cut-edge list of the same fused kernel)
        names = ["k_sensor", "k_diss_prepare", "k_edge(interior)",
                 "stage_edges: halo + boundary/cut edges | PerssonC0: RT gradient + interior edges incl. viscous flux",
                 "stage_visc: PerssonC0 boundary/cut edges incl. viscous flux", "wave", "element kernel"]
        for _ in range(2):
            for rk in range(5):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
                ev[0].record(self.stream)
                dev.stage_sensor(rk)
                if self.diss and self.world > 1 and driver != "peer":
                    self.x_vtx()
                ev[1].record(self.stream)
                dev.stage_prepare(rk)
                if self.world > 1 and driver != "peer":
                    self.x_edge()
                ev[2].record(self.stream)
                dev.stage_edges_interior(rk)
                ev[3].record(self.stream)
                dev.stage_edges(rk)
                if self.diss and self.world > 1 and driver != "peer":
                    self.x_diss()
                ev[4].record(self.stream)
                dev.stage_visc(rk)
                ev[5].record(self.stream)
                if self.world > 1:
                    if driver == "peer":
                        dev.stage_wave(rk)
                    else:
                        w = torch.as_tensor(_DevArray(dev.wavespeed_buffer(), 2), device="cuda")
                        self.dist.all_reduce(w, op=self.dist.ReduceOp.MAX)
                ev[6].record(self.stream)
                dev.stage_update(rk)
                ev[7].record(self.stream)
                evs.append(ev)
        torch.cuda.synchronize()
        return {nm: statistics.mean(e[i].elapsed_time(e[i + 1]) for e in evs) for i, nm in enumerate(names)}

    def roofline(self, ms, steps, phases):
        n, p, world = self.n, self.p, self.world
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        b_total, b_elem, b_edge = algorithmic_bytes(n)
        k_local = self.k1 - self.k0
        if self.diss:
            bv = algorithmic_bytes_visc(n)
            ach = bv * self.k_global * 5 * steps / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": "whole PerssonC0 stage (k_sensor, k_diss_prepare, k_grad_ws [DMMA], k_edge_int/k_edge with the viscous flux fused in, k_elem_ws<N,8,true>)",
                    "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / peak / world, "traffic": None,
                    "peak_source": peak_src, "bytes_per_element_stage": bv, "phase_ms": phases, "per_gpu": True}
        te = phases["element kernel"] * 1e-3
        tedge = phases["k_edge(interior)"] * 1e-3
        ach = b_elem * k_local / te / 1e9
        # per-GPU spread: every rank times its own kernels
        fr = self.reduce_list([ach / peak, -ach / peak], self.dist.ReduceOp.MAX if world > 1 else None)
        r = {"bound": "hbm",
             "kernel": ("k_elem_ws<%d,8,false,SPLIT> (warp-specialised async-copy pipeline: 4 copy warps, 8 flux+DMMA warps, 4 edge-interpolation warps)" if n >= 2 else "k_elem<%d,false>") % n,
             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": b_elem * k_local, "avg_launch_ms": te * 1e3,
             "per_gpu_frac_min_max": [-fr[1], fr[0]],
             "edge_kernel": {"achieved": b_edge * k_local / tedge / 1e9 if tedge > 0 else None, "avg_launch_ms": tedge * 1e3,
                             "note": "interior-edge kernel only; the boundary / cut-edge list runs in the next phase"},
             "phase_ms": phases,
             "whole_stage": {"bytes_per_element": b_total,
                             "achieved_per_gpu": b_total * self.k_global * 5 * steps / (ms * 1e-3) / 1e9 / world,
                             "frac": b_total * self.k_global * 5 * steps / (ms * 1e-3) / 1e9 / peak / world}}
        prof = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if world == 1 and os.path.exists(prof):
            try:
                tr = json.load(open(prof))
                r["traffic"] = tr.get("k_elem_N%d" % n)
                r["traffic_source"] = "ncu --set full dram__bytes_read+write of the same kernel, " + tr.get("source", "profiles/r02_traffic.json")
            except Exception:  # noqa: BLE001
                pass
        return r

    def e2e(self, driver, steps):
        """Through the C ABI with HOST buffers: upload the state, `steps` steps each returning step info to the host (as
        Solve does for its progress line), download the state.  The solver is device resident by design -- the seam is
        c.RK.Step, the state crosses PCIe once per run, not once per step -- so the copies are amortised over the steps;
        bytes per step are reported as total / steps."""
        torch, dev = self.torch, self.dev
        if self.world > 1:
            dev.peer_enable(driver == "peer")
        self.q_host[...] = self.c.Q
        dev.set_clock(0.0, 0)            # the run (and its checksum) starts at t = 0 whatever was timed before
        self.barrier()
        t0 = time.perf_counter()
        dev.set_state(self.q_host)
        t1 = time.perf_counter()
        info = None
        for _ in range(steps):
            if self.world == 1 or driver == "peer":
                info = dev.step(1, sync=True)
            else:
                self.nccl_step()
                info = dev.step_finish(sync=True)
        t2 = time.perf_counter()
        dev.get_state(self.q_host)
        self.barrier()
        el = self.reduce_max(time.perf_counter() - t0)
        el_loop = self.reduce_max(t2 - t1)
        qb = 4 * self.p.NpInt * (self.k1 - self.k0) * 8
        # checksum of the state after exactly `steps` steps from the initial condition: identical at every N
        own = self.q_host[:, :, self.k0 - self.k_off:self.k1 - self.k_off]
        sq = self.reduce_list([float((own[v] * own[v]).sum()) for v in range(4)], self.dist.ReduceOp.SUM if self.world > 1 else None)
        return ({"value": self.dof_per_step * steps / el, "unit": "DOF-stage-updates/s",
                 "h2d_bytes_per_step": qb / steps, "d2h_bytes_per_step": qb / steps + 40,
                 "step_loop_value": self.dof_per_step * steps / el_loop,
                 "what": "per rank: dfr2d_set_state(pinned host Q, own columns) + %d x dfr2d_step(1, info) + dfr2d_get_state; "
                         "state copies happen once per run (device-resident solver), bytes are total / steps; step_loop_value = the "
                         "%d host-synchronised steps alone, without the two state copies" % (steps, steps)},
                {"after_steps": steps, "time": info["time"], "l2": [float(np.sqrt(v)) for v in sq]})

    def close(self):
        self.dev.close()

    # ---- single process over all GPUs (driver multi_step) ----------------------------------------------------------------
    def multi_step_run(self, n_gpus, steps, warmup):
        torch, lib = self.torch, self.lib
        t0 = time.perf_counter()
        gc = self.c if self.c.window is None else build_case(self.nx, self.ny, self.n, dissipation=self.diss)
        gp = gc.problem
        build_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        devs = [lib.Dfr2d(gp, n_parts=n_gpus, part=g, device=g) for g in range(n_gpus)]
        create_s = time.perf_counter() - t0
        q_host_t = torch.empty((4, gp.NpInt, gp.K), dtype=torch.float64, pin_memory=True)
        q_host = q_host_t.numpy()
        q_host[...] = gc.Q
        for d in devs:
            d.set_state(q_host)
        lib.multi_step(devs, warmup, sync=True)

        def sync_all():
            for g in range(n_gpus):
                torch.cuda.synchronize(g)
        sync_all()
        l0 = sum(d.launch_count() for d in devs)
        ev = []
        for g in range(n_gpus):
            with torch.cuda.device(g):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(torch.cuda.default_stream(g))
                ev.append((a, b))
        t_host = time.perf_counter()
        lib.multi_step(devs, steps, sync=False)
        t_issue = time.perf_counter() - t_host
        for g in range(n_gpus):
            with torch.cuda.device(g):
                ev[g][1].record(torch.cuda.default_stream(g))
        sync_all()
        ms = max(a.elapsed_time(b) for a, b in ev)
        launches = sum(d.launch_count() for d in devs) - l0
        prof = lib.multi_step_profile(devs)                     # [n][5][phases] ms
        sync_all()
        # end to end with host buffers
        q_host[...] = gc.Q
        for d in devs:
            d.set_clock(0.0, 0)
        t0 = time.perf_counter()
        lib.multi_set_state(devs, q_host)
        t1 = time.perf_counter()
        info = None
        for _ in range(steps):
            info = lib.multi_step(devs, 1, sync=True)
        t2 = time.perf_counter()
        lib.multi_get_state(devs, q_host)
        sync_all()
        el = time.perf_counter() - t0
        l2 = [float(np.sqrt((q_host[v] * q_host[v]).sum())) for v in range(4)]
        for d in devs:
            d.close()
        torch.cuda.set_device(self.local_rank)
        per_stage = prof.sum(axis=2)                             # [n][5]
        return {"value": self.dof_per_step * steps / (ms * 1e-3), "ms_per_step": ms / steps, "gpu_launches": int(launches),
                "host_issue_ms_per_step": t_issue * 1e3 / steps, "create_s": create_s, "host_problem_build_s": build_s,
                "e2e": {"value": self.dof_per_step * steps / el, "unit": "DOF-stage-updates/s",
                        "step_loop_value": self.dof_per_step * steps / (t2 - t1),
                        "what": "one process: dfr2d_multi_set_state + %d x dfr2d_multi_step(1, info) + dfr2d_multi_get_state" % steps},
                "checksum": {"after_steps": steps, "time": info["time"], "l2": l2},
                "timeline_ms": {"phases": list(lib.PROFILE_PHASES),
                                "mean_over_partitions_and_stages": [float(v) for v in prof.mean(axis=(0, 1))],
                                "max_over_partitions_mean_over_stages": [float(v) for v in prof.mean(axis=1).max(axis=0)],
                                "stage_ms_per_partition": [[float(v) for v in row] for row in per_stage]},
                "what": "dfr2d_multi_step: ONE host thread drives all %d GPUs; halo and wave-speed exchange by peer stores + "
                        "arrival flags (csrc/dfr2d_peer.cuh); CUDA events on every device, max over devices" % n_gpus}


def workload_text(r):
    if r.diss:
        return ("%s: Sod shock tube, %dx%dx2=%d triangles, N=%d (NpInt=%d), Roe flux, global dt, PerssonC0 sensor + artificial "
                "dissipation, In/Out/Wall boundaries" % (r.workload, r.nx, r.ny, r.k_global, r.n, r.p.NpInt))
    return ("%s: isentropic vortex, %dx%dx2=%d triangles, N=%d (NpInt=%d), Roe flux, global dt, IVortex+Riemann boundaries"
            % (r.workload, r.nx, r.ny, r.k_global, r.n, r.p.NpInt))


def measure(args, workload, rank, world, local_rank, gloo, steps, warmup, full):
    """One workload -> dict.  full = the headline (clocks, cpu baseline, comparison drivers)."""
    r = Runner(args, workload, rank, world, local_rank, gloo)
    if args.driver == "multi_step":                     # single process, all GPUs, nothing else
        out = r.multi_step_run(args.gpus, steps, warmup)
        out["config"] = {"workload": workload_text(r)}
        return out
    r.open()
    driver = "peer" if (world > 1 and args.driver == "peer" and r.peer_error is None) else ("nccl" if world > 1 else "single")
    sampler = ClockSampler(local_rank)
    if rank == 0 and full:
        sampler.start()          # started before the warm-up so that samples exist even for sub-second timed regions
    ms, launches = r.timed(driver, steps, warmup)
    clocks = sampler.stop() if (rank == 0 and full) else None
    phases = r.kernel_times(driver)
    roof = r.roofline(ms, steps, phases)
    e2e, checksum = r.e2e(driver, steps)
    k_local = r.k1 - r.k0
    p = r.p
    out = {
        "value": r.dof_per_step * steps / (ms * 1e-3), "ms_per_step": ms / steps, "gpu_launches": launches,
        "element_stages_per_s": r.k_global * 5 * steps / (ms * 1e-3), "us_per_element_iteration": ms * 1e3 / steps / r.k_global,
        "driver": driver, "e2e": e2e, "checksum": checksum, "roofline": roof, "clocks": clocks,
        "config": {"workload": workload_text(r),
                   "partition": "PartitionMap.Split1D element ranges over %d GPU(s)" % world,
                   "exchange": {"single": "none (one partition)",
                                "peer": "P2P stores into the partner's mailbox + arrival flags over CUDA IPC peer memory (no NCCL, no host in the stage loop)",
                                "nccl": "host-driven stage API, NCCL all_to_all_single + all_reduce(MAX)"}[driver],
                   "l2": "no flush needed: per-GPU working set %.2f GB >> 126 MB L2"
                         % ((5 * 4 * p.NpInt + 12 * p.NpEdge + 6 * p.NpEdge) * 8 * k_local / 1e9),
                   "setup_s": {"host_problem_build": r.build_s, "dfr2d_create": r.create_s,
                               "host_problem": ("window of %d of %d elements per rank (dfr2d_create_window)" % (p.K, r.k_global))
                               if r.c.window else "global"}},
    }
    if world > 1 and r.peer_error:
        out["config"]["peer_error"] = r.peer_error
    if world > 1 and full and driver == "peer" and not args.no_compare:
        # the comparison line: round 1's NCCL-moved, Python-driven protocol on the same handles
        ms2, _ = r.timed("nccl", max(2, steps // 2), 3)
        e2e2, ck2 = r.e2e("nccl", max(2, steps // 2))
        out["nccl_driver"] = {"value": r.dof_per_step * max(2, steps // 2) / (ms2 * 1e-3), "ms_per_step": ms2 / max(2, steps // 2),
                              "e2e": e2e2["value"], "checksum": ck2}
    r.close()
    if world > 1 and full and not args.no_compare:
        # the single-process driver (what a Go host calls), on the same GPUs, after every rank has released its partition
        r.dist.barrier(group=gloo)
        if rank == 0:
            try:
                out["multi_step"] = r.multi_step_run(world, steps, warmup)
            except Exception as e:  # noqa: BLE001
                out["multi_step"] = {"error": repr(e)}
        r.dist.barrier(group=gloo)
    out["_n"] = r.n
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--driver", default="peer", choices=["peer", "nccl", "multi_step"])
    ap.add_argument("--nx", type=int, default=0, help="override the mesh size (debug)")
    ap.add_argument("--order", type=int, default=-1, help="override the polynomial order (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the C2 / C3 lines under 'also'")
    ap.add_argument("--no-compare", action="store_true", help="skip the nccl / multi_step comparison drivers at N > 1")
    ap.add_argument("--global-problem", action="store_true", help="N > 1: every rank builds the global problem (round 1 behaviour)")
    ap.add_argument("--sample-only", action="store_true", help="reference arm: only the bounded sample, not the full mesh")
    ap.add_argument("--cpu-full-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if not args.cpu_full_worker:
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gloo = None
    if world > 1:
        if args.driver == "multi_step":
            raise SystemExit("--driver multi_step is one process over all GPUs: run it without torchrun")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")
    t_all = time.perf_counter()
    head = measure(args, args.workload, rank, world, local_rank, gloo, args.steps, args.warmup, True)
    n = head.pop("_n", WORKLOADS[args.workload][2])
    also = {}
    if not args.no_also and args.nx == 0 and args.order < 0:
        for wl, k in (("c2", 200), ("c3", 10)):
            if wl == args.workload:
                continue
            m = measure(args, wl, rank, world, local_rank, gloo, k, 5 if wl == "c2" else 3, False)
            m.pop("_n", None)
            also[wl] = {"metric": "DOF-stage-updates/s", "value": m["value"], "ms_per_step": m["ms_per_step"], "steps": k,
                        "driver": m.get("driver", args.driver), "e2e": m["e2e"]["value"], "gpu_launches": m["gpu_launches"],
                        "checksum": m.get("checksum"),
                        "roofline_frac": (m["roofline"]["frac"] if m.get("roofline") and m["roofline"].get("per_gpu")
                                          else (m["roofline"]["whole_stage"]["frac"] if m.get("roofline") else None)),
                        "roofline": m.get("roofline"), "workload": m["config"]["workload"]}
    if rank == 0:
        line = {
            "metric": "DOF-stage-updates/s", "value": head["value"], "unit": "DOF-stage-updates/s",
            "n_gpus": args.gpus if args.driver == "multi_step" else world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        }
        line.update({k: v for k, v in head.items() if k not in ("value", "ms_per_step")})
        if also:
            line["also"] = also
        line["go_reference"] = go_probe()
        if not args.no_cpu_baseline and world == 1 and args.driver != "multi_step":
            line["cpu_baseline"] = cpu_baseline(n, dissipation=args.workload in DISSIPATION_WORKLOADS)
        line["wall_s"] = time.perf_counter() - t_all
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
