import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure)."""
    from oracle import binding
    return binding.load()


@pytest.fixture(scope="session")
def gpu():
    """Device + queue through the C ABI.  GPU tests never fall back to the CPU: no device => error."""
    from tests.util import GpuContext
    ctx = GpuContext()
    yield ctx
    ctx.close()
