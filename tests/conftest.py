import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def V():
    """The product binding; raises (not skips) if libvsc_b200.so is missing."""
    import vsc_b200

    vsc_b200.lib()
    return vsc_b200


@pytest.fixture(scope="session")
def O():
    from oracle import oracle

    oracle.lib()
    return oracle
