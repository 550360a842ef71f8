"""Loader of libeonb200.so -- the only native library this package ever opens.

There is no CPU path: if the CUDA library has not been built, or cannot be loaded, importing the
product API raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C eon_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _capi

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libeonb200.so")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")

_lib = None


class BackendUnavailable(RuntimeError):
    """libeonb200.so is missing or unusable (no CPU fallback exists)."""


def build(verbose: bool = False) -> str:
    """Compile eon_b200/csrc into eon_b200/libeonb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC_DIR, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise BackendUnavailable("building libeonb200.so failed:\n" + out.stdout[-4000:])
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendUnavailable(
                f"{LIB_PATH} not found: the CUDA backend has not been built (make -C eon_b200/csrc). "
                "eon_b200 has no CPU fallback.")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as exc:  # missing libcudart / driver
            raise BackendUnavailable(f"cannot load {LIB_PATH}: {exc}") from exc
        _capi.bind(lib)
        if lib.eonb200_abi_version() != _capi.ABI_VERSION:
            raise BackendUnavailable("libeonb200.so ABI version mismatch")
        _lib = lib
    return _lib
