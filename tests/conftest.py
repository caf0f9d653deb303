import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def state_dicts():
    from conan_b200 import synth
    return synth.make_all_state_dicts(1234)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
