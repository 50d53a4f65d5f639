"""pytest configuration: the `gpu` marker and shared fixtures."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import vm_oracle
    vm_oracle.lib()
    return vm_oracle


@pytest.fixture(scope="session")
def vm():
    """The product package (directory name contains a dot, so it is loaded by path)."""
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture
def rng():
    return np.random.default_rng(20240601)
