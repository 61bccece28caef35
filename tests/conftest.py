import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("oracle", "stereo-vision_b200", ""):
    path = os.path.join(ROOT, sub) if sub else ROOT
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import checkers
    checkers.build("oracle")
    return checkers.OracleElas()


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Built when /root/reference is present,
    otherwise the prebuilt .so that travelled with the snapshot; skipped if neither exists."""
    import checkers
    if os.path.isdir("/root/reference/libelas/src"):
        checkers.build("ref")
    if not checkers.have_ref():
        pytest.skip("oracle/_ref/libelas_ref.so not available")
    return checkers.RefElas()
