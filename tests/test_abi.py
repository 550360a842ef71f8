"""The C-ABI library loads without a GPU and exports every symbol include/eonb200.h declares; without a
device the product fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from eon_b200 import _capi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "eonb200.h")) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"EONB200_API[^;(]*?\b(eonb200_\w+)\s*\(", text)))


def test_header_symbols_are_declared_in_binding():
    assert header_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.eonb200_abi_version() == _capi.ABI_VERSION


def test_struct_layouts_match_defaults():
    """config_default / search_params_default run on the host; the values pin the ctypes struct layouts."""
    lib = _lib.load()
    cfg = _capi.Config()
    lib.eonb200_config_default(C.byref(cfg), _capi.POT_MORSE_PT)
    assert (cfg.kind, cfg.cutoff, cfg.morse_De, cfg.morse_a, cfg.morse_re, cfg.skin) == (2, 9.5, 0.7102, 1.6047, 2.8970, 1.0)
    assert cfg.lpa_override == 0 and cfg.tiles_override == 0
    p = _capi.SearchParams()
    lib.eonb200_search_params_default(C.byref(p))
    # include/eon/Parameters.h defaults read by this path
    assert (p.finite_difference, p.max_move, p.opt_converged_force) == (0.01, 0.2, 0.01)
    assert (p.lbfgs_memory, p.dimer_rotations_max, p.lanczos_max_iterations) == (20, 10, 20)
    assert p.saddle_max_energy == 20.0 and p.saddle_nonlocal_distance_abort == 0.0


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_create_fails_loudly_without_gpu():
    from eon_b200 import B200Potential, PotentialError
    with pytest.raises(PotentialError) as exc:
        B200Potential("eam")
    assert "no CPU fallback" in str(exc.value) or "CUDA" in str(exc.value)


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/ (the product path must not)."""
    pkg = os.path.join(ROOT, "eon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert "oracle_lib" not in text and "liboracle" not in text and "emu_lib" not in text, f
