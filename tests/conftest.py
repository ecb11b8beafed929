"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"`: oracle vs golden vectors / compiled reference, host-side logic, C-ABI exports (no GPU).
`-m gpu`      : parity tests proper, CUDA path through the C ABI vs the oracle (needs a B200).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the native pieces are built in-tree; building is a no-op when they are fresh
    import build_native
    build_native.build()
    from oracle import oracle as O
    O.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def capi():
    from fpsample_b200 import capi as C
    return C


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
