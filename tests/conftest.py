"""pytest configuration: the `gpu` marker and shared CPU-checker fixtures.

`-m "not gpu"`: oracle vs the compiled reference / golden vectors, host logic, ABI symbols.
`-m gpu`      : the parity tests proper (CUDA path through the C-ABI vs the oracle).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build(ref=os.path.isdir("/root/reference"))
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import pyoracle
    if not pyoracle.Reference.available(False):
        pytest.skip("oracle/_ref/libclover_ref.so not built (needs /root/reference)")
    return pyoracle.Reference(stochastic=False)


@pytest.fixture(scope="session")
def reference_sr():
    from oracle import pyoracle
    if not pyoracle.Reference.available(True):
        pytest.skip("oracle/_ref/libclover_ref_sr.so not built (needs /root/reference)")
    return pyoracle.Reference(stochastic=True)
