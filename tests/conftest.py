import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_once():
    so = os.path.join(ROOT, "vbsky_b200", "libvbsky_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g

        g.build()


_build_once()


@pytest.fixture(scope="session")
def cuda():
    import torch

    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    return torch.device("cuda", 0)
