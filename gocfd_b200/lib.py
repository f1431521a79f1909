"""ctypes binding of the C ABI in include/dfr2d.h (the same entry points the cgo shim binds).

There is no fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFR2D_LIB_PATH") or os.path.join(_HERE, "csrc", "libdfr2d.so")     # override: A/B of two builds

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Freestream(C.Structure):
    _fields_ = [("Gamma", C.c_double), ("Qinf", C.c_double * 4), ("Pinf", C.c_double), ("QQinf", C.c_double),
                ("Cinf", C.c_double), ("Alpha", C.c_double), ("Minf", C.c_double)]


class Vortex(C.Structure):
    _fields_ = [("Beta", C.c_double), ("X0", C.c_double), ("Y0", C.c_double), ("Gamma", C.c_double),
                ("Ufs", C.c_double)]


class ProblemStruct(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("flux_type", C.c_int32), ("init_case", C.c_int32), ("local_time_stepping", C.c_int32),
        ("max_iterations", C.c_int32), ("dissipation", C.c_int32),
        ("K", C.c_int64), ("NV", C.c_int64), ("NE", C.c_int64), ("NBP", C.c_int64),
        ("CFL", C.c_double), ("FinalTime", C.c_double), ("Kappa", C.c_double),
        ("FSFar", Freestream), ("FSIn", Freestream), ("FSOut", Freestream), ("vortex", Vortex),
        ("FluxEdgeInterp", _dp), ("DivInt", _dp), ("Div", _dp), ("V", _dp), ("Vinv", _dp),
        ("MassMatrix", _dp), ("D", _dp), ("P", _dp), ("ModeFilter", _dp), ("Bary", _dp),
        ("Jdet", _dp), ("Jinv", _dp), ("FaceNormX", _dp), ("FaceNormY", _dp), ("IInII", _dp), ("EdgeLenMax", _dp),
        ("EToV", _ip), ("edge_kL", _ip), ("edge_kR", _ip), ("edge_numL", _ip), ("edge_numR", _ip),
        ("edge_nconn", _ip), ("edge_bc", _ip), ("edge_len", _dp), ("EtoEdge", _ip),
        ("bp_edge", _ip), ("bp_x", _dp), ("bp_y", _dp),
    ]


class StepInfo(C.Structure):
    _fields_ = [("time", C.c_double), ("dt", C.c_double), ("steps", C.c_int64), ("finished", C.c_int32),
                ("nan_found", C.c_int32)]


EXPORTS = [
    "dfr2d_create", "dfr2d_destroy", "dfr2d_last_error", "dfr2d_set_state", "dfr2d_get_state", "dfr2d_step",
    "dfr2d_residual", "dfr2d_rhs", "dfr2d_set_register", "dfr2d_get_register", "dfr2d_get_field",
    "dfr2d_set_stream", "dfr2d_partition_range", "dfr2d_halo_counts", "dfr2d_halo_buffers",
    "dfr2d_wavespeed_buffer", "dfr2d_stage_prepare", "dfr2d_stage_edges", "dfr2d_stage_update",
    "dfr2d_step_finish", "dfr2d_launch_count", "dfr2d_stage_sensor", "dfr2d_stage_visc", "dfr2d_stage_edges_interior",
    "dfr2d_exchange_counts", "dfr2d_exchange_buffers", "dfr2d_plan_vertices", "dfr2d_plot_field", "dfr2d_init_state", "dfr2d_rcm_order", "dfr2d_grad_mma_table", "dfr2d_multi_step", "dfr2d_mma_diss_table",
    "dfr2d_peer_export", "dfr2d_peer_connect", "dfr2d_peer_enable", "dfr2d_multi_set_state", "dfr2d_multi_get_state",
    "dfr2d_set_clock", "dfr2d_epsilon_field", "dfr2d_create_window", "dfr2d_plan_create_window", "dfr2d_stage_wave", "dfr2d_multi_step_profile",
    "dfr2d_capture_edge_values", "dfr2d_gradient_field", "dfr2d_multi_create", "dfr2d_multi_destroy", "dfr2d_multi_residual", "dfr2d_hilbert_order",
    "dfr2d_plan_create", "dfr2d_plan_destroy", "dfr2d_plan_sizes", "dfr2d_plan_edges", "dfr2d_plan_halo",
]

_lib = None


def load():
    """Load libdfr2d.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("CUDA library %s is missing: build it with __graft_entry__.build(); "
                           "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.dfr2d_create.argtypes = [C.POINTER(ProblemStruct), C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    lib.dfr2d_create_window.argtypes = [C.POINTER(ProblemStruct), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    lib.dfr2d_plan_create_window.argtypes = [C.POINTER(ProblemStruct), C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(H)]
    lib.dfr2d_destroy.argtypes = [H]
    lib.dfr2d_destroy.restype = None
    lib.dfr2d_last_error.argtypes = [H]
    lib.dfr2d_last_error.restype = C.c_char_p
    for name in ("dfr2d_set_state", "dfr2d_get_state"):
        getattr(lib, name).argtypes = [H, _dp]
    lib.dfr2d_step.argtypes = [H, C.c_int, C.POINTER(StepInfo)]
    lib.dfr2d_residual.argtypes = [H, _dp]
    lib.dfr2d_rhs.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_set_register.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_get_register.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_get_field.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_init_state.argtypes = [H, C.c_int, C.c_int64, _dp, _dp, _ip, _dp, _dp]
    lib.dfr2d_plot_field.argtypes = [H, C.c_int, _dp, C.c_int, C.POINTER(C.c_float)]
    lib.dfr2d_set_stream.argtypes = [H, C.c_void_p]
    lib.dfr2d_partition_range.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.dfr2d_halo_counts.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.dfr2d_halo_buffers.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.dfr2d_wavespeed_buffer.argtypes = [H, C.POINTER(C.c_void_p)]
    lib.dfr2d_exchange_counts.argtypes = [H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.dfr2d_exchange_buffers.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.dfr2d_plan_vertices.argtypes = [H, C.POINTER(C.c_int64), _ip]
    for name in ("dfr2d_stage_prepare", "dfr2d_stage_edges", "dfr2d_stage_update", "dfr2d_stage_sensor", "dfr2d_stage_visc",
                 "dfr2d_stage_edges_interior"):
        getattr(lib, name).argtypes = [H, C.c_int]
    lib.dfr2d_step_finish.argtypes = [H, C.POINTER(StepInfo)]
    lib.dfr2d_launch_count.argtypes = [H]
    lib.dfr2d_launch_count.restype = C.c_int64
    lp = C.POINTER(C.c_int64)
    lib.dfr2d_plan_create.argtypes = [C.POINTER(ProblemStruct), C.c_int, C.c_int, C.POINTER(H)]
    lib.dfr2d_plan_destroy.argtypes = [H]
    lib.dfr2d_plan_destroy.restype = None
    lib.dfr2d_plan_sizes.argtypes = [H, lp]
    lib.dfr2d_plan_edges.argtypes = [H, _ip, _ip, _ip, lp, _ip]
    lib.dfr2d_plan_halo.argtypes = [H, lp, lp, lp, _ip, _ip, _ip, _ip]
    lib.dfr2d_rcm_order.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _ip, _ip]
    lib.dfr2d_hilbert_order.argtypes = [C.c_int64, C.c_int64, _ip, _dp, _dp, _ip]
    lib.dfr2d_mma_diss_table.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64]
    lib.dfr2d_mma_diss_table.restype = C.c_int64
    lib.dfr2d_multi_step.argtypes = [C.POINTER(H), C.c_int, C.c_int, C.POINTER(StepInfo)]
    lib.dfr2d_peer_export.argtypes = [H, C.c_void_p]
    lib.dfr2d_peer_connect.argtypes = [H, C.c_void_p, C.c_int]
    lib.dfr2d_stage_wave.argtypes = [H, C.c_int]
    lib.dfr2d_peer_enable.argtypes = [H, C.c_int]
    lib.dfr2d_multi_set_state.argtypes = [C.POINTER(H), C.c_int, _dp]
    lib.dfr2d_multi_create.argtypes = [C.POINTER(ProblemStruct), C.c_int, _ip, C.POINTER(H)]
    lib.dfr2d_multi_destroy.argtypes = [C.POINTER(H), C.c_int]
    lib.dfr2d_multi_destroy.restype = None
    lib.dfr2d_multi_residual.argtypes = [C.POINTER(H), C.c_int, _dp]
    lib.dfr2d_multi_get_state.argtypes = [C.POINTER(H), C.c_int, _dp]
    lib.dfr2d_set_clock.argtypes = [H, C.c_double, C.c_int64]
    lib.dfr2d_epsilon_field.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_capture_edge_values.argtypes = [H, C.c_int, _dp, _dp]
    lib.dfr2d_gradient_field.argtypes = [H, C.c_int, _dp]
    lib.dfr2d_multi_step_profile.argtypes = [C.POINTER(H), C.c_int, C.POINTER(C.c_float)]
    lib.dfr2d_grad_mma_table.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int64]
    lib.dfr2d_grad_mma_table.restype = C.c_int64
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _fs(fs):
    a = fs.as_array()
    return Freestream(a[0], (C.c_double * 4)(*a[1:5]), a[5], a[6], a[7], a[8], a[9])


def problem_struct(p):
    """Flatten a host-side Problem into the C struct.  Returns (struct, keepalive list)."""
    keep = []

    def dd(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        keep.append(x)
        return _d(x)

    def ii(x):
        x = np.ascontiguousarray(x, dtype=np.int32)
        keep.append(x)
        return _i(x)

    v = p.Vortex.as_array()
    s = ProblemStruct(
        N=p.N, flux_type=p.FluxType, init_case=p.Case, local_time_stepping=int(p.LocalTimeStepping),
        max_iterations=p.MaxIterations, dissipation=int(p.Dissipation),
        K=p.K, NV=p.NV, NE=p.NE, NBP=p.NBP, CFL=p.CFL, FinalTime=p.FinalTime, Kappa=p.Kappa,
        FSFar=_fs(p.FSFar), FSIn=_fs(p.FSIn), FSOut=_fs(p.FSOut), vortex=Vortex(*v),
        FluxEdgeInterp=dd(p.FluxEdgeInterp), DivInt=dd(p.DivInt), Div=dd(p.Div), V=dd(p.V), Vinv=dd(p.Vinv),
        MassMatrix=dd(p.MassMatrix), D=dd(p.D), P=dd(p.P), ModeFilter=dd(p.ModeFilter), Bary=dd(p.Bary),
        Jdet=dd(p.Jdet), Jinv=dd(p.Jinv), FaceNormX=dd(p.FaceNormX), FaceNormY=dd(p.FaceNormY),
        IInII=dd(p.IInII), EdgeLenMax=dd(p.EdgeLenMax),
        EToV=ii(p.EToV), edge_kL=ii(p.edge_kL), edge_kR=ii(p.edge_kR), edge_numL=ii(p.edge_numL),
        edge_numR=ii(p.edge_numR), edge_nconn=ii(p.edge_nconn), edge_bc=ii(p.edge_bc),
        edge_len=dd(p.edge_len), EtoEdge=ii(p.EtoEdge),
        bp_edge=ii(p.bp_edge), bp_x=dd(p.bp_x), bp_y=dd(p.bp_y),
    )
    return s, keep


def grad_mma_table(problem):
    """Operator table of the tensor-core gradient kernel for `problem`'s order (host-only; see include/dfr2d.h)."""
    lib = load()
    div = np.ascontiguousarray(problem.Div, dtype=np.float64)
    bary = np.ascontiguousarray(problem.Bary, dtype=np.float64)
    n = lib.dfr2d_grad_mma_table(problem.N, _d(div), _d(bary), None, 0)
    if n < 0:
        raise RuntimeError("dfr2d_grad_mma_table: bad request")
    out = np.zeros(n)
    lib.dfr2d_grad_mma_table(problem.N, _d(div), _d(bary), _d(out), n)
    return out


def mma_diss_table(problem):
    """A-fragment table [DivInt | Vinv | V] of k_elem_mma_diss for `problem`'s order (host-only)."""
    lib = load()
    ops = [np.ascontiguousarray(x, dtype=np.float64) for x in (problem.DivInt, problem.Vinv, problem.V)]
    n = lib.dfr2d_mma_diss_table(problem.N, _d(ops[0]), _d(ops[1]), _d(ops[2]), None, 0)
    if n < 0:
        raise RuntimeError("dfr2d_mma_diss_table: bad request")
    out = np.zeros(n)
    lib.dfr2d_mma_diss_table(problem.N, _d(ops[0]), _d(ops[1]), _d(ops[2]), _d(out), n)
    return out


def multi_step(devs, nsteps=1, sync=True):
    """dfr2d_multi_step over the partitions `devs` (Dfr2d objects created with n_parts=len(devs), part=g): the
    single-process multi-GPU driver -- peer copies for the halo, a peer-reading kernel for the wave-speed maximum."""
    lib = load()
    arr = (C.c_void_p * len(devs))(*[d.h for d in devs])
    info = StepInfo()
    rc = lib.dfr2d_multi_step(arr, len(devs), nsteps, C.byref(info) if sync else None)
    if rc != 0:
        raise Dfr2dError("dfr2d_multi_step failed (%d): %s" % (rc, "; ".join(lib.dfr2d_last_error(d.h).decode() for d in devs)))
    return {"time": info.time, "dt": info.dt, "steps": int(info.steps), "finished": bool(info.finished)}


def multi_set_state(devs, q):
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.float64)
    assert q.shape == devs[0].shape
    arr = (C.c_void_p * len(devs))(*[d.h for d in devs])
    if lib.dfr2d_multi_set_state(arr, len(devs), _d(q)) != 0:
        raise Dfr2dError("dfr2d_multi_set_state failed: %s" % "; ".join(lib.dfr2d_last_error(d.h).decode() for d in devs))


def multi_get_state(devs, out=None):
    lib = load()
    q = np.zeros(devs[0].shape) if out is None else out
    arr = (C.c_void_p * len(devs))(*[d.h for d in devs])
    if lib.dfr2d_multi_get_state(arr, len(devs), _d(q)) != 0:
        raise Dfr2dError("dfr2d_multi_get_state failed: %s" % "; ".join(lib.dfr2d_last_error(d.h).decode() for d in devs))
    return q


PEER_BLOB_BYTES = 1280
PROFILE_PHASES = ("sensor+prepare+pack/put", "interior edges", "halo wait + boundary/cut edges (+ RT gradient)",
                  "viscous edges", "wave put+gather", "element update")


def multi_step_profile(devs):
    """One profiled dfr2d_multi_step step: array [n][5 stages][6 phases] of milliseconds (CUDA events per partition)."""
    lib = load()
    arr = (C.c_void_p * len(devs))(*[d.h for d in devs])
    out = np.zeros((len(devs), 5, len(PROFILE_PHASES)), dtype=np.float32)
    rc = lib.dfr2d_multi_step_profile(arr, len(devs), out.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise Dfr2dError("dfr2d_multi_step_profile failed (%d): %s" % (rc, "; ".join(lib.dfr2d_last_error(d.h).decode() for d in devs)))
    return out


def rcm_order(problem):
    """order[new] = old element: reverse Cuthill-McKee over the element adjacency of `problem`'s edge table."""
    lib = load()
    order = np.zeros(problem.K, dtype=np.int32)
    kl = np.ascontiguousarray(problem.edge_kL, dtype=np.int32)
    kr = np.ascontiguousarray(problem.edge_kR, dtype=np.int32)
    nc = np.ascontiguousarray(problem.edge_nconn, dtype=np.int32)
    rc = lib.dfr2d_rcm_order(problem.K, problem.NE, _i(kl), _i(kr), _i(nc), _i(order))
    if rc != 0:
        raise Dfr2dError("dfr2d_rcm_order failed (%d): %s" % (rc, lib.dfr2d_last_error(None).decode()))
    return order


def hilbert_order(etov, vx, vy):
    """order[new] = old element along a Hilbert curve through the rank-normalised element centroids (host only)."""
    lib = load()
    ev = np.ascontiguousarray(etov, dtype=np.int32).reshape(-1, 3)
    x = np.ascontiguousarray(vx, dtype=np.float64)
    y = np.ascontiguousarray(vy, dtype=np.float64)
    order = np.zeros(ev.shape[0], dtype=np.int32)
    rc = lib.dfr2d_hilbert_order(ev.shape[0], x.shape[0], _i(ev), _d(x), _d(y), _i(order))
    if rc != 0:
        raise Dfr2dError("dfr2d_hilbert_order failed (%d): %s" % (rc, lib.dfr2d_last_error(None).decode()))
    return order


class Dfr2dError(RuntimeError):
    pass


class Dfr2d:
    """One partition of the device solver.  Same surface as the oracle's OracleSolver."""

    def __init__(self, problem, n_parts=1, part=0, device=0, window=None):
        """window = (K_global, k_offset): `problem` describes only elements [k_offset, k_offset + problem.K) of the mesh
        (dfr2d_create_window); host arrays then have the window's columns."""
        self.lib = load()
        self.p = problem
        self.shape = (4, problem.NpInt, problem.K)
        s, keep = problem_struct(problem)
        self.h = C.c_void_p()
        if window is None:
            rc = self.lib.dfr2d_create(C.byref(s), n_parts, part, device, C.byref(self.h))
        else:
            rc = self.lib.dfr2d_create_window(C.byref(s), int(window[0]), int(window[1]), n_parts, part, device, C.byref(self.h))
        del keep
        if rc != 0:
            raise Dfr2dError("dfr2d_create failed (%d): %s" % (rc, self.lib.dfr2d_last_error(None).decode()))
        self.n_parts, self.part = n_parts, part

    def _ck(self, rc):
        if rc != 0:
            raise Dfr2dError("dfr2d call failed (%d): %s" % (rc, self.lib.dfr2d_last_error(self.h).decode()))

    def close(self):
        if self.h:
            self.lib.dfr2d_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.shape
        self._ck(self.lib.dfr2d_set_state(self.h, _d(q)))

    def get_state(self, out=None):
        q = np.zeros(self.shape) if out is None else out
        self._ck(self.lib.dfr2d_get_state(self.h, _d(q)))
        return q

    def set_register(self, reg, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        self._ck(self.lib.dfr2d_set_register(self.h, reg, _d(q)))

    def get_register(self, reg):
        q = np.zeros(self.shape)
        self._ck(self.lib.dfr2d_get_register(self.h, reg, _d(q)))
        return q

    def step(self, nsteps=1, sync=True):
        info = StepInfo()
        self._ck(self.lib.dfr2d_step(self.h, nsteps, C.byref(info) if sync else None))
        return {"time": info.time, "dt": info.dt, "steps": int(info.steps), "finished": bool(info.finished)}

    def set_clock(self, time=0.0, steps=0):
        self._ck(self.lib.dfr2d_set_clock(self.h, float(time), int(steps)))

    def rhs(self, rk=0):
        out = np.zeros(self.shape)
        self._ck(self.lib.dfr2d_rhs(self.h, rk, _d(out)))
        return out

    def residual(self):
        r = np.zeros(4)
        self._ck(self.lib.dfr2d_residual(self.h, _d(r)))
        return list(r)

    def get_field(self, which):
        out = np.zeros(self.p.K)
        self._ck(self.lib.dfr2d_get_field(self.h, which, _d(out)))
        return out

    def init_state(self, init_case, vx, vy, etov, r, s):
        """InitializeSolution on the device (replaces building Q on the host + set_state)."""
        vx = np.ascontiguousarray(vx, dtype=np.float64)
        vy = np.ascontiguousarray(vy, dtype=np.float64)
        etov = np.ascontiguousarray(etov, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert etov.shape == (self.p.K, 3) and r.shape == (self.p.NpInt,) and s.shape == (self.p.NpInt,)
        self._ck(self.lib.dfr2d_init_state(self.h, int(init_case), len(vx), _d(vx), _d(vy), _i(etov), _d(r), _d(s)))

    def plot_field(self, flow_function, graph_interp, out=None):
        """GetPlotField on the device: float32 [K, NpGraph] (own rows of a multi-partition handle)."""
        gi = np.ascontiguousarray(graph_interp, dtype=np.float64)
        if out is None:
            out = np.zeros((self.p.K, gi.shape[0]), dtype=np.float32)
        self._ck(self.lib.dfr2d_plot_field(self.h, int(flow_function), _d(gi), gi.shape[0],
                                           out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def epsilon_field(self, c0=False, out=None):
        """EpsilonDissipation (c0=False) / EpsilonDissipationC0 (c0=True) plot fields: [NpFlux, K] float64."""
        if out is None:
            out = np.zeros((self.p.NpFlux, self.p.K))
        self._ck(self.lib.dfr2d_epsilon_field(self.h, int(bool(c0)), _d(out)))
        return out

    def capture_edge_values(self, on=True):
        """Keep the EdgeQValues store (Q_Face of every step's last stage) for the gradient plot fields."""
        if on and not self.p.Dissipation:
            nx = np.ascontiguousarray(self.p.FaceNormX, dtype=np.float64)
            ny = np.ascontiguousarray(self.p.FaceNormY, dtype=np.float64)
            self._ck(self.lib.dfr2d_capture_edge_values(self.h, 1, _d(nx), _d(ny)))
        else:
            self._ck(self.lib.dfr2d_capture_edge_values(self.h, int(bool(on)), None, None))

    def gradient_field(self, flow_function, out=None):
        """XGradient* (200..203) / YGradient* (300..303) plot fields (plot.go:54-77): [NpFlux, K] float64."""
        if out is None:
            out = np.zeros((self.p.NpFlux, self.p.K))
        self._ck(self.lib.dfr2d_gradient_field(self.h, int(flow_function), _d(out)))
        return out

    # ---- multi-partition plumbing -------------------------------------------------------
    def set_stream(self, stream_ptr):
        self._ck(self.lib.dfr2d_set_stream(self.h, C.c_void_p(stream_ptr)))

    def partition_range(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.dfr2d_partition_range(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def halo_counts(self):
        s = (C.c_int64 * self.n_parts)()
        r = (C.c_int64 * self.n_parts)()
        self._ck(self.lib.dfr2d_halo_counts(self.h, s, r))
        return list(s), list(r)

    def halo_buffers(self):
        s, r = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.dfr2d_halo_buffers(self.h, C.byref(s), C.byref(r)))
        return s.value, r.value

    def wavespeed_buffer(self):
        p = C.c_void_p()
        self._ck(self.lib.dfr2d_wavespeed_buffer(self.h, C.byref(p)))
        return p.value

    XCHG_EDGE, XCHG_VERTEX, XCHG_DISS = 0, 1, 2

    def exchange_counts(self, which):
        s = (C.c_int64 * self.n_parts)()
        r = (C.c_int64 * self.n_parts)()
        self._ck(self.lib.dfr2d_exchange_counts(self.h, which, s, r))
        return list(s), list(r)

    def exchange_buffers(self, which):
        s, r = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.dfr2d_exchange_buffers(self.h, which, C.byref(s), C.byref(r)))
        return s.value, r.value

    def stage_sensor(self, rk):
        self._ck(self.lib.dfr2d_stage_sensor(self.h, rk))

    def stage_edges_interior(self, rk):
        self._ck(self.lib.dfr2d_stage_edges_interior(self.h, rk))

    def stage_visc(self, rk):
        self._ck(self.lib.dfr2d_stage_visc(self.h, rk))

    def stage_prepare(self, rk):
        self._ck(self.lib.dfr2d_stage_prepare(self.h, rk))

    def stage_edges(self, rk):
        self._ck(self.lib.dfr2d_stage_edges(self.h, rk))

    def stage_update(self, rk):
        self._ck(self.lib.dfr2d_stage_update(self.h, rk))

    def stage_wave(self, rk):
        self._ck(self.lib.dfr2d_stage_wave(self.h, rk))

    def peer_export(self):
        """The mailbox description a partner needs (bytes); gather one per partition and pass the list to peer_connect."""
        blob = C.create_string_buffer(PEER_BLOB_BYTES)
        self._ck(self.lib.dfr2d_peer_export(self.h, blob))
        return blob.raw

    def peer_connect(self, blobs):
        """blobs[p] = peer_export() of partition p (own entry included).  Afterwards step() works on this partition."""
        assert len(blobs) == self.n_parts and all(len(b) == PEER_BLOB_BYTES for b in blobs)
        raw = C.create_string_buffer(b"".join(blobs), PEER_BLOB_BYTES * len(blobs))
        self._ck(self.lib.dfr2d_peer_connect(self.h, raw, len(blobs)))

    def peer_enable(self, on):
        self._ck(self.lib.dfr2d_peer_enable(self.h, int(bool(on))))

    def step_finish(self, sync=True):
        info = StepInfo()
        self._ck(self.lib.dfr2d_step_finish(self.h, C.byref(info) if sync else None))
        return {"time": info.time, "dt": info.dt, "steps": int(info.steps), "finished": bool(info.finished)}

    def launch_count(self):
        return int(self.lib.dfr2d_launch_count(self.h))


class Plan:
    """Host-only partition plan of one (n_parts, part): integer tables only, no CUDA needed."""

    def __init__(self, problem, n_parts, part, window=None):
        lib = load()
        s, keep = problem_struct(problem)
        h = C.c_void_p()
        if window is None:
            rc = lib.dfr2d_plan_create(C.byref(s), n_parts, part, C.byref(h))
        else:
            rc = lib.dfr2d_plan_create_window(C.byref(s), int(window[0]), int(window[1]), n_parts, part, C.byref(h))
        del keep
        if rc != 0:
            raise Dfr2dError("dfr2d_plan_create failed (%d): %s" % (rc, lib.dfr2d_last_error(None).decode()))
        try:
            sz = (C.c_int64 * 8)()
            lib.dfr2d_plan_sizes(h, sz)
            (self.k0, self.k1, self.G, self.Kp, self.NE, self.NEp, self.n_cut, self.NBP) = [int(x) for x in sz]
            self.K = self.k1 - self.k0
            self.kL = np.zeros(self.NE, np.int32)
            self.kR = np.zeros(self.NE, np.int32)
            self.meta = np.zeros(self.NE, np.int32)
            self.global_edge = np.zeros(self.NE, np.int64)
            self.etoe = np.zeros((3, self.Kp), np.int32)
            lp = C.POINTER(C.c_int64)
            lib.dfr2d_plan_edges(h, _i(self.kL), _i(self.kR), _i(self.meta), self.global_edge.ctypes.data_as(lp),
                                 _i(self.etoe))
            self.send_counts = np.zeros(n_parts, np.int64)
            self.recv_counts = np.zeros(n_parts, np.int64)
            self.ghost_global = np.zeros(max(self.G, 1), np.int64)
            self.send_elem = np.zeros(max(self.n_cut, 1), np.int32)
            self.send_row0 = np.zeros(max(self.n_cut, 1), np.int32)
            self.recv_col = np.zeros(max(self.n_cut, 1), np.int32)
            self.recv_row0 = np.zeros(max(self.n_cut, 1), np.int32)
            lib.dfr2d_plan_halo(h, self.send_counts.ctypes.data_as(lp), self.recv_counts.ctypes.data_as(lp),
                                self.ghost_global.ctypes.data_as(lp), _i(self.send_elem), _i(self.send_row0),
                                _i(self.recv_col), _i(self.recv_row0))
            self.vertex_counts = np.zeros(n_parts, np.int64)
            lib.dfr2d_plan_vertices(h, self.vertex_counts.ctypes.data_as(lp), None)
            self.vertex_ids = np.zeros(max(int(self.vertex_counts.sum()) // 2, 1), np.int32)
            lib.dfr2d_plan_vertices(h, self.vertex_counts.ctypes.data_as(lp), _i(self.vertex_ids))
            self.vertex_ids = self.vertex_ids[:int(self.vertex_counts.sum()) // 2]
            for name in ("ghost_global",):
                setattr(self, name, getattr(self, name)[:self.G])
            for name in ("send_elem", "send_row0", "recv_col", "recv_row0"):
                setattr(self, name, getattr(self, name)[:self.n_cut])
        finally:
            lib.dfr2d_plan_destroy(h)
