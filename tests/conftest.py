import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    import torch

    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b).clamp_min(1e-30))


@pytest.fixture(scope="session")
def matcha_sd():
    from emojivoice_b200 import synthetic
    from emojivoice_b200.config import VCTK

    return synthetic.matcha_state_dict(VCTK, seed=1234)
