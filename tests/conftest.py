import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_visible():
    """True iff the CUDA driver reports at least one device (no torch import needed)."""
    import ctypes

    for name in ("libcuda.so.1", "libcuda.so"):
        try:
            cuda = ctypes.CDLL(name)
        except OSError:
            continue
        count = ctypes.c_int(0)
        if cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(count)) == 0:
            return count.value > 0
        return False
    return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box that has nvcc but no GPU."""
    if _cuda_device_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def brian():
    """The Brian2 front-end (reference install under baseline/_ref) with the b200 device registered."""
    import brian2_b200  # noqa: F401  (registers the device, makes brian2 importable)
    import brian2

    return brian2


@pytest.fixture()
def project_dir(request):
    """Generated projects are built IN-TREE (git-ignored) so that the loaded .so files are the
    repository's own native code."""
    import re
    import shutil

    name = re.sub(r"[^A-Za-z0-9_]+", "_", request.node.name)
    path = os.path.join(ROOT, "brian2_b200", "_prebuilt", f"test_{name}")
    shutil.rmtree(path, ignore_errors=True)
    return path
