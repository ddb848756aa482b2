import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import edipack_oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def engine():
    """The product library initialised on cuda:0 (GPU tests only)."""
    import edipack_b200 as E

    E.ed_init(0)
    yield E
    E.ed_finalize()
