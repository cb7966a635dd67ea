import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "pl-slam_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds (or re-uses) the product library and the C oracle."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def fe(built):
    """A Frontend on cuda:0.  Fails loudly (no skip, no fallback) when there is no GPU."""
    import plslam_b200 as plf
    f = plf.Frontend(device=0)
    yield f
    f.close()
