import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C-ABI.  Fails loudly (no CPU path) if unusable."""
    from crnn_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
