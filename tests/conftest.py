import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def brian():
    """The Brian2 front-end (reference install under oracle/_ref) with the b200 device registered."""
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
