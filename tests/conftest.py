import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import mcp_b200
        ctx = mcp_b200.Context(0)
        ctx.close()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """A live mcp context.  GPU tests FAIL (not skip) when libmcp.so or the GPU is missing:
    the CUDA path is the product and there is no fallback to hide behind."""
    import mcp_b200
    c = mcp_b200.Context(0)
    yield c
    c.close()
