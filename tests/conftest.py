import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def engine_exact():
    import mp3rgain_b200 as m
    e = m.Engine(device=0, exact=True)
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_fast():
    import mp3rgain_b200 as m
    e = m.Engine(device=0, exact=False)
    yield e
    e.close()
