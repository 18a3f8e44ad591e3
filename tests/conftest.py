import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def hostsim():
    import support
    return support.HostSim()


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    import support
    return support.CudaBackend()
