import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The shared library must exist; tests never fall back to anything else."""
    from librempeg_b200 import build as native
    if not os.path.exists(native.SO_PATH):
        native.build_native()
    yield


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped, not failed, on a machine without a CUDA device."""
    try:
        from librempeg_b200 import swscale as S
        have = S.device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
