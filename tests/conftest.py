import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_ops():
    """The reference's own compiled COMMON_OPS (oracle/_ref), or skip when it was not built."""
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
    return mod
