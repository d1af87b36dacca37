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


@pytest.fixture(autouse=True)
def _oracle_lu_form(request):
    """GPU parity tests compare against the oracle with its two named switches ON: the CUDA kernels store the diagonal
    of U inverted and multiply in the triangular solves, and their log/exp/pow are the lean functions of
    crnn_b200/csrc/lean_math.h, which the oracle then runs from the same header (DESIGN.md "Named deviations").  The
    CPU tests keep the oracle's defaults (division LU, C library math); tests/test_full_batch_parity_gpu.py measures
    the literal forms against the kernels as well."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    from oracle import oracle
    with oracle.lu_reciprocal(True), oracle.shared_math(True), oracle.kc4_inverse(True):
        yield
