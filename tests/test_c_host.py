"""A compiled C host on the C ABI: host_c/solve_host.c drives include/dfr2d.h with the cgo shim's call sequence
(go/dfr2d/dfr2d.go; the Go toolchain is absent, so this is the closest executable stand-in for the patched Euler.Solve
loop, model_problems/Euler2D/euler.go:171-190).  No Python, ctypes or torch between the host and the library.

CPU part: the header compiles as strict C11 (-Wall -Wextra -Werror -pedantic; cgo compiles it as C), the program links
against libdfr2d.so, the struct layout of the ctypes mirror equals the header's, and without a device the run fails
loudly (no CPU fallback).  GPU part: what the C host writes equals the oracle (1e-11) and, bitwise, the ctypes host."""
import os
import subprocess

import numpy as np
import pytest

import c_host_pack
from conftest import ROOT, mesh_path
from gocfd_b200.host.euler2d import Euler
from gocfd_b200.host.input_parameters import InputParameters2D
from gocfd_b200.host.meshgen import structured_tri_mesh

HOST_DIR = os.path.join(ROOT, "host_c")
HOST = os.path.join(HOST_DIR, "solve_host")
TOL = 1e-11


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-s", "-C", HOST_DIR], check=True)
    return HOST


def case(**kw):
    base = dict(CFL=1.0, FluxType="Roe", InitType="IVortex", PolynomialOrder=2, FinalTime=50.0, MaxIterations=1000,
                Gamma=1.4, Minf=0.1)
    mesh = kw.pop("mesh", None) or structured_tri_mesh(14, 9)
    base.update(kw)
    return Euler(InputParameters2D(**base), mesh)


def run(host, c, tmp_path, nsteps, *more, check=True):
    pack, out = str(tmp_path / "problem.pack"), str(tmp_path / "result.bin")
    c_host_pack.write_pack(pack, c.problem, c.Q)
    r = subprocess.run([host, pack, out, str(nsteps)] + [str(m) for m in more], capture_output=True, text=True)
    if check:
        assert r.returncode == 0, r.stderr
        return c_host_pack.read_out(out)
    return r


def rel_l2(a, b):
    return np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b)


def test_header_is_strict_c_and_host_links(host):
    """gcc -std=c11 -pedantic -Werror accepted include/dfr2d.h and the program resolved every symbol it calls."""
    out = subprocess.run(["ldd", host], capture_output=True, text=True, check=True).stdout
    assert "libdfr2d.so" in out and "not found" not in out.split("libdfr2d.so")[1].splitlines()[0]


def test_rejects_a_file_that_is_no_pack(host, tmp_path):
    bad = tmp_path / "bad.pack"
    bad.write_bytes(b"not a pack at all")
    r = subprocess.run([host, str(bad), str(tmp_path / "o"), "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "not a problem pack" in r.stderr


def test_struct_layout_matches_and_no_device_fails_loudly(host, tmp_path):
    """The scalar section written from the ctypes mirror has exactly offsetof(dfr2d_problem, FluxEdgeInterp) bytes (else
    the host stops with 'struct layout mismatch').  On a machine without a GPU the run then dies in dfr2d_create with the
    CUDA error: there is no CPU path behind the boundary."""
    from conftest import _cuda_device_count
    r = run(host, case(), tmp_path, 1, check=False)
    assert "struct layout mismatch" not in r.stderr
    if _cuda_device_count() == 0:
        assert r.returncode != 0 and "dfr2d_create" in r.stderr
    else:
        assert r.returncode == 0, r.stderr


def test_truncated_scalar_section_is_a_layout_error(host, tmp_path):
    pack = tmp_path / "short.pack"
    c = case()
    c_host_pack.write_pack(str(pack), c.problem, c.Q)
    raw = pack.read_bytes()
    n = int.from_bytes(raw[8:16], "little")
    pack.write_bytes(raw[:8] + (n - 8).to_bytes(8, "little") + raw[16:16 + n - 8] + raw[16 + n:])
    r = subprocess.run([host, str(pack), str(tmp_path / "o"), "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "struct layout mismatch" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(PolynomialOrder=2), dict(PolynomialOrder=4, FluxType="lax"),
                                dict(PolynomialOrder=3, LocalTimeStepping=True)],
                         ids=["N2-roe", "N4-lax", "N3-localdt"])
def test_c_host_matches_oracle_and_ctypes_host(host, tmp_path, kw):
    """Vortex on 252 triangles, 7 steps with a residual read every step (PrintUpdate's cadence at its densest)."""
    from gocfd_b200 import lib
    from oracle.euler2d_oracle import OracleSolver
    c = case(**kw)
    info, maxr, q = run(host, c, tmp_path, 7)
    ora = OracleSolver(c.problem)
    ora.set_state(c.Q)
    for _ in range(7):
        ora.step()
    assert info.steps == 7 and not info.finished and not info.nan_found
    assert rel_l2(q, ora.get_state()) < TOL
    assert abs(info.time - ora.Time) <= 1e-12 * max(1.0, abs(ora.Time))
    np.testing.assert_allclose(maxr, ora.residual(), rtol=1e-8, atol=1e-13)
    dev = lib.Dfr2d(c.problem)
    dev.set_state(c.Q)
    for _ in range(7):
        dev.step(1)
    assert np.array_equal(q, dev.get_state().ravel())
    assert np.array_equal(maxr, np.asarray(dev.residual()))
    dev.close()


@pytest.mark.gpu
def test_c_host_shock_tube_with_limiter(host, tmp_path):
    """PerssonC0 path through the C host: Sod tube on the shipped 100-point mesh, N=2, 6 steps."""
    from oracle.euler2d_oracle import OracleSolver
    c = case(PolynomialOrder=2, InitType="shocktube", FluxType="roe", CFL=1.0, Limiter="persson c0", Kappa=3.0,
             mesh=mesh_path("sod-aligned-100pts.su2"))
    info, maxr, q = run(host, c, tmp_path, 6)
    ora = OracleSolver(c.problem)
    ora.set_state(c.Q)
    for _ in range(6):
        ora.step()
    assert info.steps == 6
    assert rel_l2(q, ora.get_state()) < TOL


@pytest.mark.gpu
def test_c_host_stops_at_max_iterations(host, tmp_path):
    """CheckIfFinished (euler.go:796-801) seen from the C host: asked for 9 steps with MaxIterations = 4."""
    c = case(MaxIterations=4)
    info, _, _ = run(host, c, tmp_path, 9)
    assert info.steps == 4 and info.finished


@pytest.mark.gpu
@pytest.mark.parametrize("n_parts", [2, 5])
def test_c_host_multi_partition_is_bitwise_single(host, tmp_path, n_parts):
    """The MultiSolver sequence (dfr2d_multi_set_state / dfr2d_multi_step / dfr2d_multi_get_state) from the C host,
    partitions spread over the devices present: bitwise the single-partition run, residual maxima included."""
    from conftest import _cuda_device_count
    c = case(PolynomialOrder=3)
    _, r1, q1 = run(host, c, tmp_path, 5)
    info, rn, qn = run(host, c, tmp_path, 5, n_parts, min(n_parts, _cuda_device_count()))
    assert info.steps == 5
    assert np.array_equal(q1, qn)
    assert np.array_equal(r1, rn)


@pytest.mark.gpu
@pytest.mark.parametrize("n_parts", [1, 3])
def test_c_host_calls_migrate_between_os_threads(host, tmp_path, n_parts):
    """A goroutine that is not locked to its OS thread makes successive cgo calls from different threads (SURVEY.md 8b
    "Threading"): every step issued from a freshly created pthread gives bitwise the result of the single-threaded host --
    the library selects its device (and, with several partitions, every partition's device) inside each entry point."""
    from conftest import _cuda_device_count
    c = case(PolynomialOrder=2, LocalTimeStepping=(n_parts == 1))
    nd = min(n_parts, _cuda_device_count())
    i0, r0, q0 = run(host, c, tmp_path, 6, n_parts, nd, 0)
    i1, r1, q1 = run(host, c, tmp_path, 6, n_parts, nd, 1)
    assert i0.steps == i1.steps == 6 and i0.time == i1.time
    assert np.array_equal(q0, q1) and np.array_equal(r0, r1)
