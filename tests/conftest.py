import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# a fused compute+collective kernel whose peer never arrives gives up after this long (a failed test, not a hung box)
os.environ.setdefault("MOJO_B200_GAR_TIMEOUT_S", "30")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 GPU (run with -m gpu on the B200 box)")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name), map_location="cpu", weights_only=True)


@pytest.fixture(scope="session")
def torch_backend():
    """Registers the oracle as backend "torch" (test infrastructure only)."""
    import oracle.torch_backend as tb

    return tb


def to_dev(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    return x
