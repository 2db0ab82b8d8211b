import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def v1_paths():
    from helpers import model_paths
    try:
        return model_paths("vntsr")
    except FileNotFoundError:
        pytest.skip("reference detector graph not available (neither /root/reference nor oracle/_ref)")


@pytest.fixture(scope="session")
def v2_paths():
    from helpers import model_paths
    try:
        return model_paths("tt100k")
    except FileNotFoundError:
        pytest.skip("reference tt100k graph not available")
