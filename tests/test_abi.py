"""The C-ABI library loads on a CPU-only box and exports every symbol include/pddm.h declares, and the ctypes
signatures in _lib.py cover exactly that set (no compute call is made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pddm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pddm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from probabilisticdeepdiffusionmodels_b200 import _lib
    from probabilisticdeepdiffusionmodels_b200.build import build_library
    build_library()
    lib = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/pddm.h but not exported by libpddm_b200.so"
    assert sorted(_lib.SIGNATURES) == decl, set(_lib.SIGNATURES) ^ set(decl)


def test_version_and_strerror_without_gpu():
    from probabilisticdeepdiffusionmodels_b200 import _lib
    lib = _lib.load()
    assert lib.pddm_version() >= 100
    assert lib.pddm_strerror(0) == b"ok"
    assert b"sm_100" in lib.pddm_strerror(-5)
    for rc in (-1, -2, -3, -4, -6, -99):
        assert len(lib.pddm_strerror(rc)) > 0


def test_struct_sizes_match_header_layout():
    """ctypes mirrors of the parameter structs: spot-check sizes that follow from the C declarations."""
    from probabilisticdeepdiffusionmodels_b200 import _lib as L
    p, i = ctypes.sizeof(ctypes.c_void_p), 4
    assert ctypes.sizeof(L.Tables) == 12 * p + 8
    assert ctypes.sizeof(L.PackDesc) == 3 * p + 8 * i
    assert ctypes.sizeof(L.AttnFwdParams) == 3 * p + 4 * i
    assert ctypes.sizeof(L.AttnBwdParams) == 5 * p + 4 * i + p + 8
    assert ctypes.sizeof(L.ConvParams) == 6 * p + 3 * i + 8 * i + 4 * 9 * i + i + 6 * i + p + 2 * i
    assert ctypes.sizeof(L.AdamParams) % 8 == 0 and ctypes.sizeof(L.GnBwdParams) % 8 == 0


def test_no_cpu_fallback():
    import torch
    from probabilisticdeepdiffusionmodels_b200 import functional as F
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        F.q_sample(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 1, None)
    with pytest.raises(RuntimeError):
        F.gn_silu_fwd(torch.zeros(1, 4, 4, 32), torch.ones(32), torch.zeros(32))
