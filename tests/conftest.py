import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MESHES = os.path.join(ROOT, "tests", "golden", "meshes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def mesh_path(name):
    return os.path.join(MESHES, name)


@pytest.fixture
def meshes():
    return mesh_path
