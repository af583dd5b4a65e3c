import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return Reference()


@pytest.fixture(scope="session")
def ctx():
    from quantum_geometric_tensor_b200 import api
    c = api.Context(0)
    yield c
    c.close()
