import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) libusaug.so; nvcc cross-compiles sm_100a without a GPU."""
    from ultrasoundaugmentation_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cuda_device(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test collected on a machine without CUDA")
    return torch.device("cuda", 0)
