"""TEST INFRASTRUCTURE: the device sources compiled as plain C++ (tests/emu), one thread per CTA.
Used only by the CPU-only tests to exercise the kernels' logic; never by the eon_b200 package."""
import ctypes as C
import os
import subprocess

from eon_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
_lib = None


def lib():
    global _lib
    if _lib is None:
        out = subprocess.run(["make", "-C", EMU_DIR, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if out.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + out.stdout[-4000:])
        _lib = _capi.bind(C.CDLL(os.path.join(EMU_DIR, "libeonb200_emu.so")))
    return _lib
