"""Test configuration.

`-m "not gpu"`: oracle vs goldens / reference binaries, host logic, ABI surface, and the device
sources' logic through the one-thread-per-CTA emulation build (tests/emu, test infrastructure).
`-m gpu`: the parity tests proper -- the same checks through libeonb200.so on a B200.
"""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


def _cuda_lib():
    from eon_b200 import _lib
    return _lib.load()


def _emu_lib():
    import emu_lib
    return emu_lib.lib()


@pytest.fixture(scope="session", params=[pytest.param("emu", id="emu"),
                                         pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def backend(request):
    """The C ABI under test: libeonb200.so on the GPU, or its logic-only emulation build on CPU."""
    return _cuda_lib() if request.param == "cuda" else _emu_lib()


@pytest.fixture(scope="session")
def cuda_backend():
    return _cuda_lib()
