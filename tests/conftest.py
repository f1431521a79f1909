import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MESHES = os.path.join(ROOT, "tests", "golden", "meshes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """Number of usable CUDA devices, 0 without a driver (asks the runtime the library itself links against)."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a machine without a GPU: the gpu-marked tests are skipped, not failed (the two-tier runs are
    `-m "not gpu"` here and `-m gpu` on the B200 box)."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (no usable driver / device on this machine)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def mesh_path(name):
    return os.path.join(MESHES, name)


@pytest.fixture
def meshes():
    return mesh_path
