import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if os.environ.get("RG_SIMT_EMU") == "1":
        # Development aid (tests/dev/simt_emu): run the -m gpu tests' logic without a GPU by pointing the ctypes binding at
        # the engine sources compiled against the CPU SIMT interpreter.  Only this test session is affected; the product's
        # loader (mp3rgain_b200/_lib.py) knows nothing about it, and the driver's GPU run never sets the variable.
        sys.path.insert(0, os.path.join(ROOT, "tests", "dev", "simt_emu"))
        import build_emu
        from mp3rgain_b200 import _lib
        _lib.LIB_PATH = os.environ.get("RG_SIMT_EMU_LIB") or build_emu.build()   # (an instrumented build made by hand: coverage, ...)


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
    for kv in os.environ.get("RG_TEST_OPTS", "").split(","):   # development: e.g. RG_TEST_OPTS=k4_lean=1
        if "=" in kv:
            e.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    yield e
    e.close()
