"""pytest configuration: `gpu` marker, import paths for the product shim (`femb_b200`) and the oracle."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def libfemb():
    """Build (if needed) and load libfemb.so."""
    import femb_b200  # noqa: F401  (import shim)
    from femb_b200 import _lib, build

    build.build()
    return _lib.load()
