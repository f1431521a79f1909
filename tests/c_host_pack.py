"""Writer of the problem pack host_c/solve_host.c reads, and reader of what it writes (format: see solve_host.c).

The scalar section is the ctypes image of dfr2d_problem up to its first pointer member: the C program checks its length
against offsetof(dfr2d_problem, FluxEdgeInterp), so the ctypes mirror in gocfd_b200/lib.py and the header agree or the
run fails."""
import ctypes as C
import struct

import numpy as np

from gocfd_b200 import lib

POINTER_FIELDS = [n for n, t in lib.ProblemStruct._fields_ if t in (lib._dp, lib._ip)]


def write_pack(path, problem, Q):
    s, keep = lib.problem_struct(problem)
    first = getattr(lib.ProblemStruct, POINTER_FIELDS[0]).offset
    raw = bytes(s)[:first]
    assert len(keep) == len(POINTER_FIELDS)
    it = iter(keep)          # problem_struct appends one contiguous array per pointer member, in declaration order
    with open(path, "wb") as f:
        f.write(b"DFR2DPK1")
        f.write(struct.pack("<q", len(raw)))
        f.write(raw)
        for _ in POINTER_FIELDS:
            a = next(it)
            f.write(struct.pack("<q", a.nbytes))
            f.write(a.tobytes())
        q = np.ascontiguousarray(Q, dtype=np.float64)
        f.write(struct.pack("<q", q.nbytes))
        f.write(q.tobytes())


def read_out(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"DFR2DOUT"
        info = lib.StepInfo.from_buffer_copy(f.read(C.sizeof(lib.StepInfo)))
        maxr = np.frombuffer(f.read(32), dtype=np.float64).copy()
        (n,) = struct.unpack("<q", f.read(8))
        q = np.frombuffer(f.read(8 * n), dtype=np.float64).copy()
    return info, maxr, q
