"""The C-ABI library loads and exports every symbol include/psc.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from psiclass_b200 import abi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psc.h")).read()
    return sorted(set(re.findall(r"\b(psc_[a-z_]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert declared_symbols() == sorted(abi.SYMBOLS)


def test_libpsc_exports_every_declared_symbol():
    lib = ctypes.CDLL(abi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_shim_exports_the_same_abi():
    from psiclass_b200 import testkit
    lib = ctypes.CDLL(testkit.SHIM_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # the timing struct is the one most likely to drift
    assert ctypes.sizeof(abi.psc_timing) == 8 + 8 + 8 + 8 + 8 * 12 * 3
    assert ctypes.sizeof(abi.psc_params) == 3 * 8 + 7 * 4 + 4  # padded to 8


def test_no_cpu_fallback():
    """Without a CUDA device psc_create must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = abi.load()
    ctx = ctypes.c_void_p()
    rc = lib.psc_create(ctypes.byref(ctx), 0)
    assert rc != 0 and not ctx.value
    with pytest.raises(RuntimeError):
        abi.Engine(0)


def test_product_does_not_reference_oracle():
    """Nothing under psiclass_b200/ may include, link or execute anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "psiclass_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".inl", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r'#include\s+"[^"]*oracle', txt) or "psc_oracle" in txt:
                    bad.append(f)
    assert not bad, bad
