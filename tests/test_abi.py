"""The C-ABI library loads on a CPU-only box and exports every symbol include/mpb200.h declares;
compute entry points refuse to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(mp):
    lib = mp.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libmpb200.so does not export %s" % s
    from mpb200 import _lib
    assert set(_lib.SIGNATURES) == set(syms), "ctypes SIGNATURES out of sync with include/mpb200.h"


def test_no_cpu_fallback(mp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mp.MPB200Error):
        mp.init(0)
    lib = mp.load()
    h = ctypes.c_void_p()
    assert lib.mpb200_samples_create(None, 0, 2, ctypes.byref(h)) != 0      # not initialised -> error
    assert b"mpb200_init" in lib.mpb200_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "motionplanning.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "liboracle" not in text and "mp_oracle.h" not in text, f


def test_julia_shim_and_integration_doc_bind_real_symbols():
    """every ccall symbol named in julia/MotionPlanningB200.jl and INTEGRATION.md is declared in the header"""
    syms = set(_declared_symbols())
    for rel in ("julia/MotionPlanningB200.jl", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, rel)).read()
        used = set(re.findall(r"\(:(mpb200_[a-z0-9_]+),\s*LIB\)", text))
        assert used, rel
        assert used <= syms, (rel, used - syms)
