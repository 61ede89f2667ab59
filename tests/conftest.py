import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """CPU-side checker libraries (oracle C port, host simulation) are built on demand; the CUDA
    library is built by __graft_entry__.build() and only loaded here."""
    import __graft_entry__ as g
    g.build_oracle()
    if not os.path.exists(g.LIB):
        g.build_cuda()
    yield
