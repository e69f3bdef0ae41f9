import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc  # oracle/oracle.py

    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One libvoxcuda context on cuda:0.  Fails (does not skip) when the extension or the
    device is missing: GPU tests must never pass on a fallback."""
    from voxlogica_b200 import ops

    c = ops.Context(0)
    yield c
    c.sync()
