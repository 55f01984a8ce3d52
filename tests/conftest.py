import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native pieces exist (cheap no-op when they are up to date)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "ode_b200", "libode_b200_single.so"), os.path.join(ROOT, "oracle", "liborc_single.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    yield
