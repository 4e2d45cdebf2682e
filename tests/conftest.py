import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def mp():
    """The product package (host mirror); importing it needs no GPU."""
    import mpb200
    return mpb200


@pytest.fixture(scope="session")
def gpu(mp):
    """Initialised product library on cuda:0; fails loudly when the CUDA path is unavailable."""
    mp.init(0)
    return mp


def unpack_bits(chunks, n):
    """Julia BitVector chunk layout -> bool[n]"""
    b = np.unpackbits(np.ascontiguousarray(chunks).view(np.uint8), bitorder="little")
    return b[:n].astype(bool)
