import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle module (test infrastructure; builds liboracle.so on first use)."""
    from oracle import oracle as orc

    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def core():
    """The CUDA-backed product module mirroring `bourse.core` (GPU tests only)."""
    import bourse_b200.core as core_mod

    return core_mod
