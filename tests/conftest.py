import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library (nvcc cross-compiles on CPU) and the oracle once per session."""
    from simka_b200 import build as b
    b.build()
    from oracle import oracle
    oracle.build()
    if os.environ.get("SIMKA_B200_EMU") == "1":
        # developer switch: run the `-m gpu` tests against the kernel-logic emulator (tests/emu) on a machine without a GPU
        from emu.emu_lib import use_emulator
        use_emulator()
