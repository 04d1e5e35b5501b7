"""Test fixtures.  `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI exports, and the kernel sources
run through the CPU emulation build (tests/emu).  `-m gpu`: the parity tests proper, through libmocap_b200.so."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def emu_engine():
    """CaptureEngine over the CPU emulation build of the kernel sources (test infrastructure, see tests/emu)."""
    sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
    import build_emu
    from emu_engine import EmuEngine
    return EmuEngine(build_emu.build())


@pytest.fixture(scope="session")
def gpu_engine():
    from mocapv2_b200 import build
    from mocapv2_b200.engine import CaptureEngine
    build.build()
    return CaptureEngine()


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def engine(request):
    """Every parity test runs twice: on the emulation build here (CPU) and on the real library on the GPU box."""
    return request.getfixturevalue("emu_engine" if request.param == "emu" else "gpu_engine")
