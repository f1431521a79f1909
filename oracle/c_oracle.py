"""ORACLE -- test infrastructure, not product code.

ctypes wrapper of oracle/c/liboracle.so, the C/OpenMP restatement of the stage (inviscid and PerssonC0 paths)
(oracle/c/euler2d_stage.c).  Same surface as OracleSolver / gocfd_b200.lib.Dfr2d.  Only tests/,
__graft_entry__ and bench.py's CPU-baseline legs may import this.  The problem struct layout is
the one of include/dfr2d.h, so the flattening helper of the ctypes binding is reused.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gocfd_b200.lib import ProblemStruct, StepInfo, problem_struct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "liboracle.so")
_dp = C.POINTER(C.c_double)
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        H = C.c_void_p
        lib.ora_create.argtypes = [C.POINTER(ProblemStruct)]
        lib.ora_create.restype = H
        lib.ora_destroy.argtypes = [H]
        lib.ora_destroy.restype = None
        lib.ora_threads.restype = C.c_int
        lib.ora_set_threads.argtypes = [C.c_int]
        lib.ora_set_threads.restype = C.c_int
        lib.ora_set_state.argtypes = [H, _dp]
        lib.ora_get_state.argtypes = [H, _dp]
        lib.ora_residual.argtypes = [H, _dp]
        lib.ora_step.argtypes = [H, C.c_int, C.POINTER(StepInfo)]
        lib.ora_rhs.argtypes = [H, C.c_int, _dp]
        lib.ora_set_register.argtypes = [H, C.c_int, _dp]
        _lib = lib
    return _lib


def threads():
    return int(load().ora_threads())


def set_threads(n):
    """Use n OpenMP threads regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1); returns the count in force."""
    return int(load().ora_set_threads(int(n)))


class COracleSolver:
    def __init__(self, problem):
        self.lib = load()
        self.p = problem
        self.shape = (4, problem.NpInt, problem.K)
        s, keep = problem_struct(problem)
        self.h = self.lib.ora_create(C.byref(s))
        del keep
        if not self.h:
            raise ValueError("ora_create failed")

    def close(self):
        if self.h:
            self.lib.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.shape
        self.lib.ora_set_state(self.h, q.ctypes.data_as(_dp))

    def set_register(self, reg, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.shape
        self.lib.ora_set_register(self.h, reg, q.ctypes.data_as(_dp))

    def get_state(self):
        q = np.zeros(self.shape)
        self.lib.ora_get_state(self.h, q.ctypes.data_as(_dp))
        return q

    def residual(self):
        r = np.zeros(4)
        self.lib.ora_residual(self.h, r.ctypes.data_as(_dp))
        return list(r)

    def rhs(self, rk=0):
        out = np.zeros(self.shape)
        self.lib.ora_rhs(self.h, rk, out.ctypes.data_as(_dp))
        return out

    def step(self, nsteps=1):
        info = StepInfo()
        self.lib.ora_step(self.h, nsteps, C.byref(info))
        return {"time": info.time, "dt": info.dt, "steps": int(info.steps), "finished": bool(info.finished)}
