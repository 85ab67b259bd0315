import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def b():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import b200arrays
    b200arrays._lib.load()  # fail loudly if the extension is missing on a GPU box
    assert b200arrays._lib.load().b200_check_device() == 0, "not an sm_100 device"
    return b200arrays
